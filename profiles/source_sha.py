"""sha256 of the CUDA sources of the library (csrc/*.cuh + atmrt_lib.cu): the key that ties an ncu capture under profiles/
to the code it was taken from. bench.py prints `roofline.frac` only when profiles/ncu_summary.json carries the sha of
the sources it runs.  usage: python profiles/source_sha.py"""
import glob
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_sha():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "atm_raytracer_b200", "csrc")
    for f in sorted(glob.glob(os.path.join(d, "*.cuh")) + [os.path.join(d, "atmrt_lib.cu")]):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(source_sha())
