"""Hygiene of the oracle / product boundary: the product never touches oracle/, and nothing that runs
on the GPU box reads /root/reference."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _files(top, exts):
    for d, _, fs in os.walk(os.path.join(ROOT, top)):
        if "__pycache__" in d:
            continue
        for f in fs:
            if f.endswith(exts):
                yield os.path.join(d, f)


def test_product_never_imports_or_links_the_oracle():
    pat = re.compile(r'(^\s*(from|import)\s+oracle\b)|liboracle|#include\s*"[^"]*oracle|oracle_render|dlopen[^\n]*oracle', re.M)
    for path in list(_files("atm_raytracer_b200", (".py", ".cu", ".cuh", ".cpp", ".h", "Makefile"))) + list(_files("include", (".h",))):
        text = open(path, errors="ignore").read()
        assert not pat.search(text), f"{path} references the oracle"


def test_nothing_reads_the_reference_checkout_at_run_time():
    for top in ("atm_raytracer_b200", "oracle", "tests"):
        for path in _files(top, (".py", ".cu", ".cuh", ".cpp")):
            if path.endswith("test_boundary.py"):
                continue
            text = open(path, errors="ignore").read()
            for m in re.finditer(r"/root/reference", text):
                line = text[text.rfind("\n", 0, m.start()) + 1:text.find("\n", m.end())]
                assert line.lstrip().startswith(("//", "#", "*", '"""')) or "paths relative to" in line, f"{path}: {line}"
    for f in ("bench.py", "__graft_entry__.py"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert "/root/reference" not in open(p).read()
