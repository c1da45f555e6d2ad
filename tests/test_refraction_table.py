"""The ray-path stage's g(h) table (atmrt_refraction_table) against the f64 oracle,
without a GPU: the table is built on the host.

g(h) = dn/n with dn the reference's central difference over +-0.01 m. The oracle's f64 evaluation of it
carries a relative noise of ~3e-7 per evaluation (rounding of n = 1 + 2.8e-4), so the pointwise check is
statistical (unbiased to a few 1e-9 over 20000 altitudes), and the sharp check is the integral
identity  int g dh = ln n(h2) - ln n(h1), which pins the table against the oracle's n(h) to ~1e-12.
"""
import numpy as np
import pytest

from atm_raytracer_b200 import runtime
from conftest import scene


def _custom(a):
    """A surface inversion (thin temperature functions inside one table cell), an isothermal function
    and two lapse rates; humid."""
    a.humidity = 0.6
    a.pressure_altitude, a.pressure = 0.0, 100800.0
    a.temperature_altitude, a.temperature = 0.0, 288.0
    grads = [-0.0065, 0.08, 0.0, -0.0065, 0.0]
    starts = [0.0, 1830.0, 1880.0, 2500.0, 11000.0]
    a.n_functions = len(grads)
    for i, (g, s) in enumerate(zip(grads, starts)):
        a.fn_gradient[i], a.fn_start_altitude[i] = g, s


class Table:
    def __init__(self, atmosphere, wavelength):
        self.cells, self.base, self.ch, self.served, self.npieces = runtime.refraction_table(atmosphere, wavelength)

    def cell(self, h):
        return np.rint((np.asarray(h, float) - self.base) / self.ch).astype(int)

    def g(self, h):
        j = self.cell(h)
        u = 2.0 * (h - (self.base + j * self.ch)) / self.ch
        c = self.cells[:, j]
        return sum(c[q] * u**q for q in range(self.cells.shape[0]))

    def integral(self, h1, h2):
        tot = 0.0
        for j in range(int(self.cell(h1)), int(self.cell(h2)) + 1):
            centre = self.base + j * self.ch
            lo, hi = max(h1, centre - self.ch / 2), min(h2, centre + self.ch / 2)
            ul, uh = 2 * (lo - centre) / self.ch, 2 * (hi - centre) / self.ch
            tot += sum(self.cells[q, j] * (uh ** (q + 1) - ul ** (q + 1)) / (q + 1) for q in range(self.cells.shape[0])) * self.ch / 2
        return tot


def _oracle_g(oracle_lib, p, h, eps=0.01):
    n0 = oracle_lib.atmosphere(p.atmosphere, p.wavelength, h)[2]
    n1 = oracle_lib.atmosphere(p.atmosphere, p.wavelength, h - eps)[2]
    n2 = oracle_lib.atmosphere(p.atmosphere, p.wavelength, h + eps)[2]
    return (n2 - n1) / (2 * eps) / n0


@pytest.mark.parametrize("custom", [False, True])
def test_table_is_the_oracles_function(oracle_lib, custom):
    p, _, _, _ = scene("c2", 0.04)
    if custom:
        _custom(p.atmosphere)
    t = Table(p.atmosphere, p.wavelength)
    served = ~np.isnan(t.cells[0])
    assert t.served == served.sum()
    # everything below the altitude where the last temperature function reaches 1 K is served, except the
    # cells that hold the start of a temperature function: those are served by pieces, two or more each
    starts = [p.atmosphere.fn_start_altitude[i] for i in range(1, p.atmosphere.n_functions)]
    boundary_cells = set(int(t.cell(s)) for s in starts)
    low = np.arange(1, int(t.cell(80000.0)))
    assert set(low[~served[low]]) <= boundary_cells | set(c + d for c in boundary_cells for d in (-1, 1))
    assert t.npieces >= 2 * len(set(low[~served[low]]))
    rng = np.random.default_rng(1)
    h = rng.uniform(-400.0, 60000.0, 40000)
    h = h[served[t.cell(h)]]
    assert h.size > 30000
    rel = t.g(h) / _oracle_g(oracle_lib, p, h) - 1.0
    low_air = h < 12000.0  # n - 1 shrinks with altitude and the oracle's rounding noise grows like 1 / (n - 1)
    assert np.abs(rel[low_air]).max() < 5e-6
    assert abs(rel[low_air].mean()) < 2e-8, rel[low_air].mean()  # unbiased: noise / sqrt(N)
    # the integral identity, inside single temperature functions and across served cells only
    spans = [(0.0, 1500.0), (2700.0, 10800.0), (12000.0, 19000.0)] if custom else [(-300.0, 10800.0), (12000.0, 19000.0), (21000.0, 31000.0)]
    for h1, h2 in spans:
        n = oracle_lib.atmosphere(p.atmosphere, p.wavelength, np.array([h1, h2]))[2]
        want = np.log1p(n[1] - 1.0) - np.log1p(n[0] - 1.0)
        assert abs(t.integral(h1, h2) / want - 1.0) < 1e-11, (h1, h2)


def test_table_rejects_what_it_cannot_serve():
    p, _, _, _ = scene("c2", 0.04)
    t = Table(p.atmosphere, p.wavelength)
    # the edge cells catch NaN / out-of-range altitudes; US-76's last function reaches 0 K at ~178 km
    assert np.isnan(t.cells[:, 0]).all() and np.isnan(t.cells[:, -1]).all()
    assert np.isnan(t.cells[:, int(t.cell(180000.0))]).all()
    assert not np.isnan(t.cells[:, int(t.cell(170000.0))]).any()
    assert 690 <= t.served <= 766
