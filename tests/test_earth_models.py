"""Known answers for the earth models beyond Spherical / FlatDistorted, on the CPU
oracle: EllipsoidCalc is Vincenty's direct formula (directional_calc.rs:88-185), checked on the published
Flinders Peak -> Buninyong line and against spherical limits; AzEqCalc is a straight line in the
azimuthal-equidistant plane (:9-28); ObserverAe walks a sphere of proj_radius but lives in the flat world
(earth_model/mod.rs:127-131, 37-51, 80-91); to_shape (mod.rs:95-112) decides the ray physics."""
import math

import numpy as np
import pytest

from atm_raytracer_b200 import abi, config
from conftest import scene

DEG = 1.0e7 / 90.0


def dms(d, m, s):
    return d + m / 60.0 + s / 3600.0


def test_vincenty_direct_flinders_peak_to_buninyong(oracle_lib):
    # Geoscience Australia's worked example on GRS80 (Vincenty 1975): 54 972.271 m at azimuth 306 52' 05.37"
    a, f = 6378137.0, 1.0 / 298.257222101
    b = a * (1.0 - f)
    lat1, lon1 = -dms(37, 57, 3.72030), dms(144, 25, 29.52440)
    az = dms(306, 52, 5.37)
    lat, lon = oracle_lib.coords_at_dist(abi.EARTH_ELLIPSOID, a, lat1, lon1, az, np.array([54972.271, 0.0]), ellipsoid_b=b)
    assert abs(lat[0] - (-dms(37, 39, 10.15610))) < 2e-8  # 2 mm
    assert abs(lon[0] - dms(143, 55, 35.38390)) < 2e-8
    assert abs(lat[1] - lat1) < 1e-12 and abs(lon[1] - lon1) < 1e-12


def test_ellipsoid_with_equal_axes_is_the_sphere(oracle_lib):
    r = 6371000.0
    d = np.array([-15.0, 15.0, 1000.0, 250000.0])
    for az in (0.0, 37.0, 90.0, 200.0):
        la, lo = oracle_lib.coords_at_dist(abi.EARTH_ELLIPSOID, r, 45.3, 6.2, az, d, ellipsoid_b=r)
        ls, los = oracle_lib.coords_at_dist(abi.EARTH_SPHERICAL, r, 45.3, 6.2, az, d)
        np.testing.assert_allclose(la, ls, atol=1e-11)
        np.testing.assert_allclose(lo, los, atol=1e-11)
    # as_cartesian: N = a for a sphere
    np.testing.assert_allclose(oracle_lib.as_cartesian(abi.EARTH_ELLIPSOID, r, 45.3, 6.2, 1200.0, ellipsoid_b=r),
                               oracle_lib.as_cartesian(abi.EARTH_SPHERICAL, r, 45.3, 6.2, 1200.0), rtol=1e-15)


def test_wgs84_cartesian_known_points(oracle_lib):
    a, b = config.WGS84_A, config.WGS84_B
    np.testing.assert_allclose(oracle_lib.as_cartesian(abi.EARTH_ELLIPSOID, a, 0.0, 0.0, 0.0, ellipsoid_b=b), [a, 0.0, 0.0], atol=1e-9)
    np.testing.assert_allclose(oracle_lib.as_cartesian(abi.EARTH_ELLIPSOID, a, 90.0, 0.0, 100.0, ellipsoid_b=b), [0.0, 0.0, b + 100.0], atol=1e-6)
    x, y, z = oracle_lib.as_cartesian(abi.EARTH_ELLIPSOID, a, 45.0, 90.0, 0.0, ellipsoid_b=b)
    assert abs(x) < 1e-9 and abs((y / a) ** 2 + (z / b) ** 2 - 1.0) < 1e-14  # on the ellipsoid


def test_azimuthal_equidistant_walks_straight_lines(oracle_lib):
    d = np.array([0.0, 1000.0, 50000.0])
    lat, lon = oracle_lib.coords_at_dist(abi.EARTH_AZIMUTHAL_EQUIDISTANT, 0.0, 45.0, 6.0, 0.0, d)  # north: towards the pole
    np.testing.assert_allclose(lat, 45.0 + d / DEG, atol=1e-12)
    np.testing.assert_allclose(lon, 6.0, atol=1e-12)
    lat, lon = oracle_lib.coords_at_dist(abi.EARTH_AZIMUTHAL_EQUIDISTANT, 0.0, 45.0, 6.0, 90.0, d)  # east: a tangent, it leaves the parallel
    r0 = 45.0 * DEG
    np.testing.assert_allclose(lat, 90.0 - np.hypot(r0, d) / DEG, atol=1e-11)
    np.testing.assert_allclose(lon, 6.0 + np.degrees(np.arctan2(d, r0)), atol=1e-11)
    # the world of the flat family: north points at the pole in the plane, up is z
    n, e, u = oracle_lib.world_directions(abi.EARTH_AZIMUTHAL_EQUIDISTANT, 0.0, 45.0, 90.0)
    np.testing.assert_allclose(n, [0.0, -1.0, 0.0], atol=1e-15)
    np.testing.assert_allclose(e, [-1.0, 0.0, 0.0], atol=1e-15)
    np.testing.assert_allclose(u, [0.0, 0.0, 1.0])
    np.testing.assert_allclose(oracle_lib.as_cartesian(abi.EARTH_AZIMUTHAL_EQUIDISTANT, 0.0, 45.0, 90.0, 300.0), [0.0, r0, 300.0], atol=1e-6)


def test_observer_ae_walks_a_sphere_in_a_flat_world(oracle_lib):
    d = np.array([100.0, 30000.0])
    la, lo = oracle_lib.coords_at_dist(abi.EARTH_OBSERVER_AE, 6371000.0, 45.0, 6.0, 33.0, d)
    ls, los = oracle_lib.coords_at_dist(abi.EARTH_SPHERICAL, 6371000.0, 45.0, 6.0, 33.0, d)
    np.testing.assert_array_equal(la, ls)
    np.testing.assert_array_equal(lo, los)
    np.testing.assert_array_equal(oracle_lib.as_cartesian(abi.EARTH_OBSERVER_AE, 6371000.0, 45.0, 6.0, 10.0),
                                  oracle_lib.as_cartesian(abi.EARTH_FLAT_DISTORTED, 0.0, 45.0, 6.0, 10.0))


@pytest.mark.parametrize("name,flat,radius", [("c3_wgs84", False, (2 * config.WGS84_A + config.WGS84_B) / 3), ("c3_azeq", True, None),
                                               ("c3_obsae", True, None), ("c3_simple", False, 6371000.0)])
def test_to_shape_decides_the_ray_physics(oracle_lib, name, flat, radius):
    """A straight ray is the shape's closed form: h0 + x tan(theta) on a plane, r0 cos(theta) / cos(phi + theta) - R on a sphere."""
    p, terrain, _, _ = scene(name, 0.02)
    p.straight_rays = 1
    w = oracle_lib.path_cache(p, terrain.tiles, 0)
    theta = math.radians(p.tilt + 0.5 * p.fov * p.height / p.width * (1.0 if p.height % 2 == 0 else (p.height - 1) / p.height))
    x, h0 = w["dist"], w["elev"][0]
    if flat:
        want = h0 + x * math.tan(theta)
    else:
        want = (radius + h0) * math.cos(theta) / np.cos(x / radius + theta) - radius
    np.testing.assert_allclose(w["elev"], want, rtol=0, atol=2e-3)  # theta of row 0 to the pixel-centring convention
    # the two shapes differ by the curvature drop x^2 / 2R = 3 km at 200 km: the test tells them apart
    assert abs((h0 + x[-1] * math.tan(theta)) - ((6371000.0 + h0) * math.cos(theta) / math.cos(x[-1] / 6371000.0 + theta) - 6371000.0)) > 2000.0
