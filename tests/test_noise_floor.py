"""The noise floor of the reference's ray integration, measured on the oracle alone.

dn/dh is a central difference over +-0.01 m of n = 1 + x with x ~ 2.8e-4: the rounding of `1 + x`
(half an ulp of 1.0 = 1.1e-16) is a relative 2e-7 of n(h+eps) - n(h-eps) ~ 5e-10, an independent error
on every one of the 12 evaluations of every RK4 step (atm-refraction 0.6 as restated in
oracle/atmrt_oracle.cpp). Two correct f64 implementations of the same formulas
(different libm, different association) therefore produce paths that differ like two realisations of
that noise. This test measures the floor without any GPU: moving the observer by ONE NANOMETRE -- far
below every tolerance of the north star -- moves the path by ~1e-6 m at 200 km, three orders more than
the perturbation. tests/test_gpu_parity.py sizes its path-altitude tolerance (5e-6 m) from this.
"""
import numpy as np

from atm_raytracer_b200 import abi
from conftest import scene

PATH_ATOL = 5e-6  # metres; used by the GPU parity tests


def test_one_nanometre_moves_the_path_by_a_micrometre(oracle_lib):
    worst = 0.0
    for name in ("c2", "c3_flat"):
        p, terrain, _, _ = scene(name, 0.05)
        p.altitude.kind, p.altitude.value = abi.ALT_ABSOLUTE, 1800.0
        q = abi.Params.from_buffer_copy(p)
        q.altitude.value = 1800.0 + 1e-9
        for y in (0, p.height // 2):
            a = oracle_lib.path_cache(p, terrain.tiles, y)
            b = oracle_lib.path_cache(q, terrain.tiles, y)
            n = min(len(a["elev"]), len(b["elev"]))
            assert n > 3900
            d = np.abs(a["elev"][:n] - b["elev"][:n])
            worst = max(worst, float(d.max()))
            # the deviation grows with distance like an integrated random walk, it is not the 1e-9 offset
            assert d[: n // 10].max() < d.max()
    assert 1e-7 < worst < PATH_ATOL, worst
