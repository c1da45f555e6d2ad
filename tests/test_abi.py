"""The C-ABI library loads on a CPU-only box, exports every symbol include/atmrt.h declares, agrees
with the ctypes mirror on struct sizes, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from atm_raytracer_b200 import abi, runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(atmrt_(?:host_)?[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    names = _declared("include/atmrt.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(runtime.lib, n), f"libatmrt_cuda.so does not export {n}"
    assert set(names) == set(runtime.EXPORTED)


def test_host_library_exports_every_declared_symbol():
    from atm_raytracer_b200 import host

    for n in _declared("atm_raytracer_b200/csrc/host/atmrt_host.h"):
        assert hasattr(host.lib, n), f"libatmrt_host.so does not export {n}"


def test_struct_sizes_match_the_header():
    out = (C.c_size_t * 16)()
    n = runtime.lib.atmrt_abi_sizes(out, 16)
    mirror = [abi.Altitude, abi.AtmosphereDef, abi.Params, abi.TileDesc, abi.Object, abi.Meta, abi.TracePoint, abi.Stats, abi.StageMs, abi.KernelMs]
    assert n == len(mirror)
    assert [out[i] for i in range(n)] == [C.sizeof(m) for m in mirror]
    assert runtime.lib.atmrt_abi_version() == 2


def test_packed_terrain_size_needs_no_gpu():
    import numpy as np

    t = runtime.Terrain.from_arrays([(45, 5, np.zeros((1201, 1201), np.int16)), (46, 6, np.zeros((1201, 1201), np.int16))])
    descs, _, n = t.c_arrays()
    b = C.c_size_t()
    assert runtime.lib.atmrt_terrain_packed_bytes(descs, n, C.byref(b)) == 0
    # 8x8 micro-tiles: 151 x 151 x 64 posts x 2 B per tile + tile table + lookup
    assert 2 * 151 * 151 * 128 <= b.value <= 2 * 151 * 151 * 128 + 4096


def _no_gpu():
    import torch

    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_create_fails_loudly_without_a_gpu():
    with pytest.raises(runtime.AtmrtError) as e:
        runtime.Context(0)
    assert e.value.code in (-3, -2)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
