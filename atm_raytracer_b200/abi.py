"""ctypes mirror of ``include/atmrt.h`` (the C ABI of the CUDA library).

Field order and types must match the header exactly; ``tests/test_abi.py`` checks the sizes against
``atmrt_abi_sizes()`` exported by the library.
"""
import ctypes as C

MAX_ATM_FUNCTIONS = 16
MAX_SPLINE_POINTS = 64
FUNCTION_LINEAR, FUNCTION_SPLINE = 0, 1
SPLINE_NATURAL, SPLINE_DERIVATIVES, SPLINE_SECOND_DERIVATIVES = 0, 1, 2
MAX_OBJECTS = 64
MAX_STEP_POINTS = 16

EARTH_SPHERICAL, EARTH_FLAT_DISTORTED, EARTH_ELLIPSOID, EARTH_AZIMUTHAL_EQUIDISTANT, EARTH_OBSERVER_AE = 0, 1, 2, 3, 4
FLAT_FAMILY = (EARTH_FLAT_DISTORTED, EARTH_AZIMUTHAL_EQUIDISTANT, EARTH_OBSERVER_AE)  # the world is the azimuthal-equidistant plane
ALT_ABSOLUTE, ALT_RELATIVE = 0, 1
GENERATOR_FAST, GENERATOR_RECTILINEAR, GENERATOR_INTERPOLATING_RECTILINEAR = 0, 1, 2
COLORING_SIMPLE, COLORING_SHADING = 0, 1
PALETTE_LEGACY, PALETTE_IMPROVED = 0, 1
OBJECT_FRUSTUM, OBJECT_BILLBOARD = 0, 1


class Altitude(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("value", C.c_double)]


class AtmosphereDef(C.Structure):
    _fields_ = [
        ("pressure_altitude", C.c_double),
        ("pressure", C.c_double),
        ("temperature_altitude", C.c_double),
        ("temperature", C.c_double),
        ("humidity", C.c_double),
        ("n_functions", C.c_int32),
        ("n_spline_points", C.c_int32),
        ("fn_start_altitude", C.c_double * MAX_ATM_FUNCTIONS),
        ("fn_gradient", C.c_double * MAX_ATM_FUNCTIONS),
        ("fn_kind", C.c_int32 * MAX_ATM_FUNCTIONS),
        ("fn_boundary", C.c_int32 * MAX_ATM_FUNCTIONS),
        ("fn_boundary_values", (C.c_double * 2) * MAX_ATM_FUNCTIONS),
        ("fn_first_point", C.c_int32 * MAX_ATM_FUNCTIONS),
        ("fn_n_points", C.c_int32 * MAX_ATM_FUNCTIONS),
        ("spline_points", (C.c_double * 2) * MAX_SPLINE_POINTS),
    ]


class Params(C.Structure):
    _fields_ = [
        ("latitude", C.c_double),
        ("longitude", C.c_double),
        ("altitude", Altitude),
        ("direction", C.c_double),
        ("tilt", C.c_double),
        ("fov", C.c_double),
        ("max_distance", C.c_double),
        ("earth_model", C.c_int32),
        ("straight_rays", C.c_int32),
        ("radius", C.c_double),
        ("ellipsoid_b", C.c_double),
        ("wavelength", C.c_double),
        ("simulation_step", C.c_double),
        ("atmosphere", AtmosphereDef),
        ("terrain_alpha", C.c_double),
        ("coloring", C.c_int32),
        ("palette", C.c_int32),
        ("water_level", C.c_double),
        ("ambient_light", C.c_double),
        ("light_dir", C.c_double * 3),
        ("simple_max_distance", C.c_double),
        ("fog_enabled", C.c_int32),
        ("generator", C.c_int32),
        ("fog_distance", C.c_double),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("x0", C.c_int32),
        ("x1", C.c_int32),
    ]


class TileDesc(C.Structure):
    _fields_ = [
        ("lat0", C.c_int32),
        ("lon0", C.c_int32),
        ("nlon", C.c_int32),
        ("nlat", C.c_int32),
        ("min_lat", C.c_double),
        ("min_lon", C.c_double),
        ("lat_interval", C.c_double),
        ("lon_interval", C.c_double),
    ]


class Object(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("texture_width", C.c_int32),
        ("texture_height", C.c_int32),
        ("_pad", C.c_int32),
        ("latitude", C.c_double),
        ("longitude", C.c_double),
        ("altitude", Altitude),
        ("r1", C.c_double),
        ("r2", C.c_double),
        ("width", C.c_double),
        ("height", C.c_double),
        ("color", C.c_double * 4),
    ]


class Meta(C.Structure):
    _fields_ = [("lat", C.c_double), ("lon", C.c_double), ("elevation", C.c_double), ("distance", C.c_double)]


class TracePoint(C.Structure):
    _fields_ = [
        ("lat", C.c_double),
        ("lon", C.c_double),
        ("distance", C.c_double),
        ("elevation", C.c_double),
        ("path_length", C.c_double),
        ("normal", C.c_double * 3),
        ("color", C.c_double * 4),
        ("is_terrain", C.c_int32),
        ("step", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("ray_steps", C.c_uint64),
        ("trace_points", C.c_uint64),
        ("pixels_hit", C.c_uint64),
        ("step_overflows", C.c_uint64),
        ("terrain_samples", C.c_uint64),
        ("path_steps", C.c_uint64),
        ("n_terrain", C.c_int32),
        ("n_path_max", C.c_int32),
        ("ms_terrain", C.c_float),
        ("ms_paths", C.c_float),
        ("ms_march", C.c_float),
        ("ms_total", C.c_float),
        ("kernel_launches", C.c_int32),
        ("_pad", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("_")}


class StageMs(C.Structure):
    _fields_ = [("ms_terrain", C.c_double), ("ms_paths", C.c_double), ("ms_march", C.c_double), ("ms_total", C.c_double),
                ("renders", C.c_int32), ("_pad", C.c_int32)]


KERNEL_COUNT = 7


class KernelMs(C.Structure):
    _fields_ = [("ms", C.c_double * KERNEL_COUNT), ("renders_with", C.c_int32 * KERNEL_COUNT), ("renders", C.c_int32), ("_pad", C.c_int32)]


def us_76() -> AtmosphereDef:
    """``AtmosphereDef::us_76()`` of the external atm-refraction crate (params.rs:453): US Standard
    Atmosphere 1976 temperature layers up to 84.852 km, sea-level fixed points 288.15 K / 101325 Pa."""
    a = AtmosphereDef()
    a.pressure_altitude = 0.0
    a.pressure = 101325.0
    a.temperature_altitude = 0.0
    a.temperature = 288.15
    a.humidity = 0.0
    starts = [0.0, 11000.0, 20000.0, 32000.0, 47000.0, 51000.0, 71000.0]
    grads = [-0.0065, 0.0, 0.001, 0.0028, 0.0, -0.0028, -0.002]
    a.n_functions = len(grads)
    for i, (s, g) in enumerate(zip(starts, grads)):
        a.fn_start_altitude[i] = s
        a.fn_gradient[i] = g
    return a
