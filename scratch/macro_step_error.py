"""Truncation error of RK4 macro steps against 25 m single steps, in x87 extended precision with the g(h) table
(no rounding noise hides it). Reproduces the figures of DESIGN.md section 4.B (macro steps).
usage: python scratch/macro_step_error.py"""
import sys, math, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import scene
from atm_raytracer_b200 import runtime
p, _, _, _ = scene("c5", 0.01)
cells, base, ch, served, npieces = runtime.refraction_table(p.atmosphere, p.wavelength)
LD = np.longdouble
cells = cells.astype(LD)
def g(h):
    j = int(round(float((h - base) / ch)))
    u = LD(2) * (h - (LD(base) + j * LD(ch))) / LD(ch)
    c = cells[:, j]
    r = LD(0)
    for q in range(6, -1, -1): r = r * u + c[q]
    return r
R = LD(6371000.0)
def f(r, v):
    return g(r - R) * (v * v + r * r) + LD(2) * v * v / r + r
def rk4(r, v, d):
    k1r, k1v = v, f(r, v)
    k2r, k2v = v + d/2*k1v, f(r + d/2*k1r, v + d/2*k1v)
    k3r, k3v = v + d/2*k2v, f(r + d/2*k2r, v + d/2*k2v)
    k4r, k4v = v + d*k3v, f(r + d*k3r, v + d*k3v)
    return r + d/6*(k1r + 2*k2r + 2*k3r + k4r), v + d/6*(k1v + 2*k2v + 2*k3v + k4v)
def run(alt, ang, step, total):
    r = R + LD(alt); v = r * LD(math.tan(math.radians(ang)))
    d = LD(step) / R
    out = []
    n = int(total / step)
    for i in range(n):
        r, v = rk4(r, v, d)
        out.append(float(r - R))
        if out[-1] < -1000 or out[-1] > 170000: break
    return np.array(out)
for ang in (-0.5, 0.0, 0.5, 1.5, 3.0, 10.0, 45.0):  # rays that cross the starts of temperature functions: step-size dependent at the mm level
    a = run(2500.0, ang, 25.0, 400000.0)
    for m in (2, 4, 8):
        b = run(2500.0, ang, 25.0 * m, 400000.0)
        n = min(len(b), len(a) // m)
        diff = np.abs(a[m-1::m][:n] - b[:n])
        print(f"ang {ang:5.1f} m {m}: steps {n}  max |h25 - h{25*m}| = {diff.max():.3e} m (final h {a[min(len(a)-1, n*m-1)]:.1f})")
print("---- smooth-zone only (below 11 km)")
for ang, alt, total in ((-0.5, 2500.0, 400000.0), (0.3, 100.0, 150000.0), (20.0, 100.0, 25000.0), (44.0, 100.0, 10000.0), (-30.0, 9000.0, 15000.0)):
    a = run(alt, ang, 25.0, total)
    for m in (8, 16, 32):
        b = run(alt, ang, 25.0 * m, total)
        n = min(len(b), len(a) // m)
        diff = np.abs(a[m-1::m][:n] - b[:n])
        print(f"ang {ang:5.1f} m {m}: steps {n}  max |h25 - h{25*m}| = {diff.max():.3e} m (final h {a[min(len(a)-1, n*m-1)]:.1f})")
