/*
 * atmrt.h -- C ABI of the B200-native panorama ray march (the drop-in boundary).
 *
 * The reference (fizyk20/atm-raytracer, Rust) has no FFI; its seam for this path is
 *   trait Generator { fn generate(&self) -> Vec<Vec<ResultPixel>> }   generator/generators/mod.rs:82-84
 *   FastGenerator::new(&Params, &Terrain, SystemTime)                 generator/generators/fast.rs:101-109
 *   renderer::draw_image(&pixels, &params) -> ImageBuffer             renderer/mod.rs:385-414
 * A thin host (`build.rs` + `extern "C"` block in Rust, or the C++ host in this repo) lowers
 * `Params` to the flat PODs below and calls these entry points; see INTEGRATION.md for the
 * reference-side binding.
 *
 * Conventions: every call returns 0 on success and a negative atmrt_status on failure; the
 * message for the last failure is available through atmrt_last_error(). The caller owns every
 * buffer it passes in; the library owns all device memory it allocates. One host thread per
 * context. No exceptions cross the ABI. All angles are degrees and all lengths metres, exactly
 * as in the reference's `Params` (generator/params.rs:496-505).
 */
#ifndef ATMRT_H
#define ATMRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATMRT_ABI_VERSION 2
#define ATMRT_MAX_ATM_FUNCTIONS 16 /* temperature functions in an atmosphere definition */
#define ATMRT_MAX_OBJECTS 64       /* objects_close is a 64-bit mask per terrain sample  */
#define ATMRT_MAX_STEP_POINTS 16   /* trace points produced by ONE march step (overflow is counted) */

typedef enum atmrt_status {
    ATMRT_OK = 0,
    ATMRT_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
    ATMRT_ERR_CUDA = -2,        /* a CUDA runtime call failed               */
    ATMRT_ERR_NO_DEVICE = -3,   /* no usable sm_100 device                  */
    ATMRT_ERR_STATE = -4,       /* call order violated (e.g. render before set_terrain) */
    ATMRT_ERR_IO = -5           /* host file IO (host helpers only)         */
} atmrt_status;

/* EarthModel (utils/earth_model/mod.rs:19-28). The parameterless variants are lowered by the host:
 * SimpleSphere = Spherical{6371000}, Wgs84 = Ellipsoid{6378137, 6356752.314245},
 * SimpleObserverAe = ObserverAe{6371000} (mod.rs:14-16, 66-74, 127-143). */
typedef enum atmrt_earth_model {
    ATMRT_EARTH_SPHERICAL = 0,              /* Spherical{radius}                                        */
    ATMRT_EARTH_FLAT_DISTORTED = 1,         /* FlatDistorted (--flat)                                   */
    ATMRT_EARTH_ELLIPSOID = 2,              /* Ellipsoid{a, b}: radius = a, ellipsoid_b = b             */
    ATMRT_EARTH_AZIMUTHAL_EQUIDISTANT = 3,  /* AzimuthalEquidistant                                     */
    ATMRT_EARTH_OBSERVER_AE = 4             /* ObserverAe{proj_radius}: radius = proj_radius            */
} atmrt_earth_model;

/* GeneratorDef (generator/mod.rs:72-78): Fast (generators/fast.rs: separable caches) or Rectilinear
 * (generators/rectilinear.rs: one ray and one azimuth walk per pixel of a rectilinear projection). */
typedef enum atmrt_generator { ATMRT_GENERATOR_FAST = 0, ATMRT_GENERATOR_RECTILINEAR = 1, ATMRT_GENERATOR_INTERPOLATING_RECTILINEAR = 2 } atmrt_generator;

/* Altitude (generator/params.rs:17-30) */
typedef enum atmrt_altitude_kind { ATMRT_ALT_ABSOLUTE = 0, ATMRT_ALT_RELATIVE = 1 } atmrt_altitude_kind;
typedef struct atmrt_altitude {
    int32_t kind; /* atmrt_altitude_kind */
    int32_t _pad;
    double value; /* metres ASL (absolute) or above terrain (relative) */
} atmrt_altitude;

/* Coloring (generator/params.rs:215-227) after ConfColoring::into_coloring (:229-267). */
typedef enum atmrt_coloring_kind { ATMRT_COLORING_SIMPLE = 0, ATMRT_COLORING_SHADING = 1 } atmrt_coloring_kind;
typedef enum atmrt_palette { ATMRT_PALETTE_LEGACY = 0, ATMRT_PALETTE_IMPROVED = 1 } atmrt_palette;

/* AtmosphereDef of the external `atm-refraction` crate as the reference's YAML exposes it
 * (README.md:281-323). Function 0 is `first_temperature_function` (valid from -infinity);
 * function i>0 starts at fn_start_altitude[i]. A function is `Linear{gradient}` or
 * `Spline{boundary_condition, points}`: a cubic spline through its (altitude, temperature) points with a
 * Natural / Derivatives[a, b] / SecondDerivatives[a, b] boundary condition, continued by its end cubics where its
 * altitude range reaches beyond the points. Temperatures: a Spline is absolute; a Linear function is continuous with
 * its neighbour towards the nearest Spline, or -- when every function is Linear -- towards the temperature fixed
 * point (ignored when a Spline is present, as the README prescribes). Pressure is hydrostatic from the pressure fixed
 * point; inside a Spline function the integral of dh / T is taken with an 8-point Gauss-Legendre rule per spline
 * segment (the crate's own quadrature is not known here). */
#define ATMRT_MAX_SPLINE_POINTS 64 /* points of all Spline functions of one atmosphere together */
typedef enum atmrt_function_kind { ATMRT_FUNCTION_LINEAR = 0, ATMRT_FUNCTION_SPLINE = 1 } atmrt_function_kind;
typedef enum atmrt_spline_boundary { ATMRT_SPLINE_NATURAL = 0, ATMRT_SPLINE_DERIVATIVES = 1, ATMRT_SPLINE_SECOND_DERIVATIVES = 2 } atmrt_spline_boundary;
typedef struct atmrt_atmosphere_def {
    double pressure_altitude; /* pressure fixed point */
    double pressure;          /* Pa */
    double temperature_altitude; /* temperature fixed point (required when all functions are Linear) */
    double temperature;          /* K */
    double humidity;             /* relative humidity 0..1, constant with altitude (default 0) */
    int32_t n_functions;         /* >= 1 */
    int32_t n_spline_points;     /* entries of spline_points in use */
    double fn_start_altitude[ATMRT_MAX_ATM_FUNCTIONS]; /* [0] unused */
    double fn_gradient[ATMRT_MAX_ATM_FUNCTIONS];       /* Linear: K per metre */
    int32_t fn_kind[ATMRT_MAX_ATM_FUNCTIONS];          /* atmrt_function_kind */
    int32_t fn_boundary[ATMRT_MAX_ATM_FUNCTIONS];      /* Spline: atmrt_spline_boundary */
    double fn_boundary_values[ATMRT_MAX_ATM_FUNCTIONS][2]; /* Spline: [a, b] of Derivatives / SecondDerivatives */
    int32_t fn_first_point[ATMRT_MAX_ATM_FUNCTIONS];   /* Spline: its points are spline_points[first .. first + n) */
    int32_t fn_n_points[ATMRT_MAX_ATM_FUNCTIONS];      /* Spline: >= 2, altitudes strictly increasing */
    double spline_points[ATMRT_MAX_SPLINE_POINTS][2];  /* (altitude m, temperature K) */
} atmrt_atmosphere_def;

/* Flat POD image of the reference's `Params` (generator/params.rs:496-505). */
typedef struct atmrt_params {
    /* view.position (params.rs:32-60) */
    double latitude, longitude;
    atmrt_altitude altitude;
    /* view.frame (params.rs:144-162) */
    double direction, tilt, fov, max_distance;
    /* model / env (params.rs:512-528; earth_model/mod.rs:95-112) */
    int32_t earth_model; /* atmrt_earth_model */
    int32_t straight_rays;
    double radius;      /* Spherical: radius; Ellipsoid: a; ObserverAe: proj_radius */
    double ellipsoid_b; /* Ellipsoid: b */
    double wavelength; /* metres; default 530e-9 */
    double simulation_step;
    atmrt_atmosphere_def atmosphere;
    /* scene */
    double terrain_alpha;
    /* view.coloring (already lowered: light_dir is the unit vector of params.rs:247-259) */
    int32_t coloring; /* atmrt_coloring_kind */
    int32_t palette;  /* atmrt_palette */
    double water_level;
    double ambient_light;
    double light_dir[3];
    double simple_max_distance; /* Coloring::Simple.max_distance = frame.max_distance */
    /* view.fog_distance: Option<f64> */
    int32_t fog_enabled;
    int32_t generator; /* atmrt_generator: output.generator (params.rs:387-392) */
    double fog_distance;
    /* output (params.rs:394-413); x0..x1 is the column block this context renders
     * (x0 = 0, x1 = width for a single GPU). */
    int32_t width, height;
    int32_t x0, x1;
} atmrt_params;

/* One DTED tile decoded on the host (terrain/mod.rs:85-98; external crate dted 0.2).
 * posts are [lon line][lat point], west->east, south->north, exactly as stored in the file. */
typedef struct atmrt_tile_desc {
    int32_t lat0, lon0;      /* HashMap key: (origin_lat as i16, origin_lon as i16), terrain/mod.rs:91-92 */
    int32_t nlon, nlat;      /* longitude lines, latitude points */
    double min_lat, min_lon; /* header origin in degrees */
    double lat_interval, lon_interval; /* arc-seconds between posts (header value / 10) */
} atmrt_tile_desc;

/* Scene object after ConfObject::into_serializable_object (object/mod.rs:165-186) except that a
 * Relative altitude is still resolved by the library (it needs Terrain::get_elev). */
typedef enum atmrt_object_kind { ATMRT_OBJECT_FRUSTUM = 0, ATMRT_OBJECT_BILLBOARD = 1 } atmrt_object_kind;
typedef struct atmrt_object {
    int32_t kind; /* atmrt_object_kind */
    int32_t texture_width, texture_height; /* billboard only; RGBA8 row-major, top row first */
    int32_t _pad;
    double latitude, longitude;
    atmrt_altitude altitude;
    double r1, r2;   /* frustum: Cylinder r1=r2, Cone r2=0 (object/mod.rs:41-54) */
    double width;    /* billboard */
    double height;   /* both */
    double color[4]; /* r,g,b,a in 0..1 (frustum) */
} atmrt_object;

/* Per-pixel metadata: the first trace point of ResultPixel.trace_points
 * (generators/mod.rs:14-30). All NaN when the ray hit nothing. */
typedef struct atmrt_meta {
    double lat, lon, elevation, distance;
} atmrt_meta;

/* One TracePoint (generators/mod.rs:21-30) + its resolved colour class. */
typedef struct atmrt_trace_point {
    double lat, lon, distance, elevation, path_length;
    double normal[3];
    double color[4]; /* Rgba(color); for Terrain(alpha): {0,0,0,alpha} */
    int32_t is_terrain;
    int32_t step; /* zip index k of the step that produced it */
} atmrt_trace_point;

typedef struct atmrt_stats {
    uint64_t ray_steps;      /* sum over pixels of zip iterations consumed by get_single_pixel */
    uint64_t trace_points;   /* total trace points produced */
    uint64_t pixels_hit;     /* pixels with >= 1 trace point */
    uint64_t step_overflows; /* steps that produced more than ATMRT_MAX_STEP_POINTS points */
    uint64_t terrain_samples; /* W_local * N_t */
    uint64_t path_steps;      /* sum over rows of stepper steps taken */
    int32_t n_terrain;       /* N_t: terrain samples per column */
    int32_t n_path_max;      /* longest path cache row (elements) */
    float ms_terrain;        /* stage A device time */
    float ms_paths;          /* stage B device time */
    float ms_march;          /* stage C (+pyramids) device time */
    float ms_total;          /* first launch -> last kernel end, device time */
    int32_t kernel_launches; /* kernels launched by this render */
    int32_t _pad;
} atmrt_stats;

/* Device time per stage, averaged over the renders since the previous atmrt_stage_times() call. */
typedef struct atmrt_stage_ms {
    double ms_terrain; /* stage A kernels (column setup, terrain profile, terrain pyramid), on their stream */
    double ms_paths;   /* stage B kernels (ray paths, path pyramid), on their stream (overlaps stage A) */
    double ms_march;   /* stage C kernel */
    double ms_total;   /* first launch -> last kernel end */
    int32_t renders;   /* how many renders were averaged */
    int32_t _pad;
} atmrt_stage_ms;

/* Device time of the hot kernels, averaged over the renders since the previous atmrt_stage_times() call (CUDA events on
 * the stream each kernel is launched on). A kernel a render did not launch does not count for it. */
enum {
    ATMRT_KERNEL_TERRAIN_PROFILE = 0, /* stage A */
    ATMRT_KERNEL_RAY_PATHS = 1,       /* stage B (refracted rays, macro steps) */
    ATMRT_KERNEL_SWEEP = 2,           /* stage C, opaque terrain without objects: the horizon sweep */
    ATMRT_KERNEL_HIT_NORMALS = 3,     /* ... normals of the distinct hit samples */
    ATMRT_KERNEL_SHADE = 4,           /* ... shading of the pixels (all row bands) */
    ATMRT_KERNEL_MARCH = 5,           /* stage C, general march */
    ATMRT_KERNEL_RECTILINEAR = 6,     /* the Rectilinear generator */
    ATMRT_KERNEL_COUNT = 7
};
typedef struct atmrt_kernel_ms {
    double ms[ATMRT_KERNEL_COUNT];
    int32_t renders_with[ATMRT_KERNEL_COUNT]; /* renders that launched the kernel */
    int32_t renders;
    int32_t _pad;
} atmrt_kernel_ms;

typedef struct atmrt_ctx atmrt_ctx;

/* ---- lifecycle -------------------------------------------------------------------------- */
int atmrt_abi_version(void);
/* sizeof() of the ABI structs in declaration order (altitude, atmosphere_def, params, tile_desc,
 * object, meta, trace_point, stats, stage_ms, kernel_ms); returns how many there are. For binding self-checks. */
int atmrt_abi_sizes(size_t* out, int n);
int atmrt_create(int device, atmrt_ctx** out);
void atmrt_destroy(atmrt_ctx* ctx);
const char* atmrt_last_error(const atmrt_ctx* ctx); /* ctx may be NULL: last create() error */

/* ---- terrain (replaces Terrain::from_folder + lazy tile load, terrain/mod.rs:35-53,66-83) -- */
/* Size of the device-side packed terrain (tile table + micro-tiled i16 posts). */
int atmrt_terrain_packed_bytes(const atmrt_tile_desc* tiles, int ntiles, size_t* bytes);
/* Upload host posts and retile them on the device into dev_dst (caller-owned device memory of
 * atmrt_terrain_packed_bytes() bytes, e.g. a torch tensor that is then NCCL-broadcast). */
int atmrt_pack_terrain(atmrt_ctx* ctx, const atmrt_tile_desc* tiles, int ntiles,
                       const int16_t* const* posts, void* dev_dst);
/* Use a packed terrain that already lives in device memory (not copied, not owned). */
int atmrt_bind_terrain(atmrt_ctx* ctx, const atmrt_tile_desc* tiles, int ntiles, const void* dev_packed);
/* Convenience: allocate + pack + bind (library-owned). */
int atmrt_set_terrain(atmrt_ctx* ctx, const atmrt_tile_desc* tiles, int ntiles,
                      const int16_t* const* posts);
/* Terrain::get_elev (terrain/mod.rs:120-126) evaluated on the device for n points; missing
 * coverage gives NaN (the reference's None). */
int atmrt_get_elev(atmrt_ctx* ctx, const double* lat, const double* lon, int n, double* elev);
/* Read back the decoded grid of one tile from the device layout ([lon][lat]); used to prove the
 * tiled layout is a bit-exact permutation of the decoded DTED posts. */
int atmrt_read_tile(atmrt_ctx* ctx, int tile_index, int16_t* posts);

/* ---- scene ------------------------------------------------------------------------------ */
int atmrt_set_params(atmrt_ctx* ctx, const atmrt_params* params);
int atmrt_set_objects(atmrt_ctx* ctx, const atmrt_object* objects, int nobjects,
                      const uint8_t* const* rgba_textures);

/* ---- render (replaces FastGenerator::generate + renderer::draw_image) -------------------- */
/* Host buffers (any may be NULL): rgb[H][x1-x0][3], meta[H][x1-x0], steps[H][x1-x0] = zip
 * iterations consumed per pixel. Copies are part of the call. The rows leave the device in bands while the sweep is still
 * working on the rows above them; that overlap needs PAGE-LOCKED buffers (atmrt_host_alloc, or cudaHostRegister'ed memory):
 * into pageable memory every band copy is host-synchronous and the call is as correct, but serial. */
int atmrt_render(atmrt_ctx* ctx, uint8_t* rgb, atmrt_meta* meta, int32_t* steps, atmrt_stats* stats);
/* Same, writing to caller-owned DEVICE buffers on `stream` (a cudaStream_t, may be NULL);
 * asynchronous unless stats != NULL (stats requires the stage timings, so it synchronises).
 * Every render of a context works in the context's caches, flags and counters: the library orders a render behind the
 * previous atmrt_render_device of the same context even when the two are given different streams (an event wait), so
 * renders of ONE context never overlap -- use one context per concurrent render. Still one host thread per context. */
int atmrt_render_device(atmrt_ctx* ctx, void* rgb_dev, void* meta_dev, void* steps_dev,
                        atmrt_stats* stats, void* stream);
/* Harvest the per-stage CUDA-event timings of every render issued since the last call (the renders
 * themselves stay asynchronous); synchronises the device. */
int atmrt_stage_times(atmrt_ctx* ctx, atmrt_stage_ms* out);
/* ResultPixel.elevation_angle and ResultPixel.azimuth (generators/mod.rs:14-19) of every pixel of the column
 * block, degrees, host buffers [H][x1-x0] each (either may be NULL). Fast generator (fast.rs:67-76):
 * get_ray_elev(y), and get_ray_dir(x) wrapped once into [0, 360); Rectilinear generator (rectilinear.rs:78-116):
 * the pixel's own elevation / direction, not wrapped. Needs set_params only. */
int atmrt_pixel_angles(atmrt_ctx* ctx, double* elevation_angle, double* azimuth);
/* Per-kernel device times of the same renders (does not reset the averages; call it before atmrt_stage_times). */
int atmrt_kernel_times(atmrt_ctx* ctx, atmrt_kernel_ms* out);
const char* atmrt_kernel_name(int i); /* "k_terrain_profile", ... for i in [0, ATMRT_KERNEL_COUNT) */
/* Full trace-point lists (ResultPixel.trace_points) for small images: points[H][x1-x0][max_points],
 * counts[H][x1-x0] (true count, may exceed max_points). Host buffers. */
int atmrt_render_trace(atmrt_ctx* ctx, atmrt_trace_point* points, int32_t* counts, int max_points);
/* 0 = default: horizon sweep for opaque terrain without objects (with the hierarchical march as its
 * device-side fallback), hierarchical min/max march otherwise; 1 = brute-force march that visits every
 * step like the reference loop (validation); 2 = hierarchical march always (validation of the sweep).
 * All three produce identical images. */
int atmrt_set_march_mode(atmrt_ctx* ctx, int mode);
/* Horizon sweep: the number of row bands a column is walked in (0 = automatic: enough bands that a narrow column
 * block still fills the GPU, clamped to bands of at least 128 rows -- or, for a block of a multi-GPU frame, two bands
 * split below the horizon, the lower one swept while the long rays above it are still being integrated; -1 = that
 * split whenever the frame allows it). Every value gives the same image (tuning and validation hook). */
int atmrt_set_sweep_bands(atmrt_ctx* ctx, int bands);
/* Ray-path stage: 0 (default) g(h) from the table, macro steps of 8 steps where g is smooth and the
 * reference's single steps across the starts of the temperature functions; 1 every evaluation through
 * libm, op for op the oracle's arithmetic, single steps (validation); 2 the table, single steps only
 * (validation of the macro steps). */
int atmrt_set_path_mode(atmrt_ctx* ctx, int mode);

/* ---- probes (the reference's TSV dumpers: elev_profile.rs:43-64, ray_path.rs:65-103,
 *      atm_printer.rs:35-46) ------------------------------------------------------------ */
/* Stage-A cache of local column x (gen_terrain_cache, utils.rs:176-199): n = N_t entries. */
int atmrt_get_terrain_profile(atmrt_ctx* ctx, int x, int capacity, double* lat, double* lon,
                              double* elev, double* normal /*[n][3]*/, uint64_t* objects_close, int* n);
/* Stage-B cache of row y (gen_path_cache, utils.rs:136-174). */
int atmrt_get_path(atmrt_ctx* ctx, int y, int capacity, double* dist, double* elev,
                   double* path_length, int* n);
/* The stepper behind `output-ray-paths` (ray_path.rs:65-91): n rays cast from altitude start_h at the given elevation
 * angles (degrees), ALWAYS refracted (cast_ray_stepper(h, ang, false)) on the shape the configured earth model lowers
 * to (to_shape), step size ray_step, nsteps calls of next(): x[nsteps] (RayState::x, the same for every ray) and
 * h[n][nsteps] (RayState::h). The ray-path stage's own step (g(h) table, pieces, libm), or with atmrt_set_path_mode(1)
 * PathStepper::next op for op. Host buffers; needs set_params only. */
int atmrt_ray_paths(atmrt_ctx* ctx, double start_h, const double* angles_deg, int n, double ray_step, int nsteps,
                    double* x, double* h);
/* The sampler behind `output-elev-profile` (elev_profile.rs:43-60): coords_at_dist_calc((lat0, lon0), azimuth)
 * .coords_at_dist(dist[i]) and Terrain::get_elev(..).unwrap_or(0.0). lat / lon may be NULL. */
int atmrt_elev_profile(atmrt_ctx* ctx, double azimuth, const double* dist, int n, double* lat, double* lon,
                       double* elev);
/* Atmosphere::temperature / pressure and Environment::n at n altitudes, on the device. */
int atmrt_atmosphere_probe(atmrt_ctx* ctx, const double* h, int n, double* temperature,
                           double* pressure, double* refractive_index);
/* The ray-path stage's atmosphere function g(h) = dn/n, dn = (n(h+eps) - n(h-eps)) / (2 eps), eps = 0.01 m
 * (atm-refraction's Environment::n / dn as the stepper combines them): as the stage
 * evaluates it without libm -- the per-atmosphere polynomial table, plus (with_pieces != 0) the pieces that
 * serve cells holding the start of a temperature function; NaN where neither serves the altitude, the
 * stage then uses libm -- and through libm. cells_served: cells the table serves. Outputs may be NULL. */
int atmrt_refraction_probe(atmrt_ctx* ctx, const double* h, int n, int with_pieces, double* g_table,
                           double* g_libm, int* cells_served);
/* The table itself, built on the host (no GPU needed): cells[ncoef][ncells], coefficient q of cell j at
 * cells[q * ncells + j]; cell j is centred on base + j * cell_height and g(h) = sum_q c_q u^q with
 * u = 2 (h - centre) / cell_height. NaN coefficients: cell not served. cells may be NULL (sizes only). */
int atmrt_refraction_table(const atmrt_atmosphere_def* def, double wavelength, double* cells, int capacity,
                           int* ncells, int* ncoef, double* base, double* cell_height, int* cells_served,
                           int* npieces);
/* Observer altitude after Altitude::abs (params.rs:23-30). */
int atmrt_observer_altitude(atmrt_ctx* ctx, double* alt);
/* FP64 FMA throughput micro-benchmark (roofline denominator): returns GFLOP/s (2 flop per FMA). */
int atmrt_fp64_peak(atmrt_ctx* ctx, double* gflops, double* dadd_ginstr);

/* ---- one panorama over the GPUs of a box, in one process (generator/mod.rs:47-99 on several GPUs) ----------------
 * The panorama shards by contiguous column blocks [i W / n, (i + 1) W / n), one context per GPU and one host thread per
 * context. group_set_terrain uploads and retiles a slice of the tiles on every GPU and all-gathers the slices over the
 * peer links (NVLink); group_render writes every GPU's block straight into the row-major host buffers of the FULL image
 * (rgb[H][W][3], meta[H][W], steps[H][W]; any may be NULL), each GPU over its own PCIe link -- allocate them with
 * atmrt_host_alloc for asynchronous copies. x0 / x1 of the params are ignored. */
typedef struct atmrt_group atmrt_group;
int atmrt_group_create(const int* devices /* NULL: 0..n-1 */, int n, atmrt_group** out);
void atmrt_group_destroy(atmrt_group* g);
const char* atmrt_group_last_error(const atmrt_group* g); /* g may be NULL: last create() error */
int atmrt_group_size(const atmrt_group* g);
/* Context i of the group (owned by the group; NULL when out of range): for the per-context probes -- atmrt_observer_altitude,
 * atmrt_atmosphere_probe, atmrt_get_path ... -- after a group render. Do not set params / terrain / objects through it. */
atmrt_ctx* atmrt_group_context(atmrt_group* g, int i);
int atmrt_group_column_block(const atmrt_group* g, int width, int i, int* x0, int* x1);
int atmrt_group_set_terrain(atmrt_group* g, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts);
int atmrt_group_set_params(atmrt_group* g, const atmrt_params* params);
int atmrt_group_set_objects(atmrt_group* g, const atmrt_object* objects, int nobjects, const uint8_t* const* rgba_textures);
int atmrt_group_render(atmrt_group* g, uint8_t* rgb, atmrt_meta* meta, int32_t* steps, atmrt_stats* stats);
/* What `gen` does between the decoded tiles and the image, as one call: atmrt_group_set_terrain + atmrt_group_render. The ray
 * paths do not read the terrain (nor does the scene preparation when every altitude is Absolute), so they are integrated
 * while the tiles are still on their way. Returns when the image is in host memory; the posts may be released then. */
int atmrt_group_render_tiles(atmrt_group* g, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts, uint8_t* rgb, atmrt_meta* meta,
                             int32_t* steps, atmrt_stats* stats);

/* atmrt_render_trace of every column block, assembled: points[H][W][max_points], counts[H][W] (true counts). Host buffers.
 * What ResultPixel.trace_points holds in the reference (generators/mod.rs:18-30). */
int atmrt_group_render_trace(atmrt_group* g, atmrt_trace_point* points, int32_t* counts, int max_points);

/* atmrt_pixel_angles of the whole image: [H][W] each, either may be NULL. */
int atmrt_group_pixel_angles(atmrt_group* g, double* elevation_angle, double* azimuth);
/* Page-locked host memory visible to every GPU (the host image / metadata / decoded tiles). */
void* atmrt_host_alloc(size_t bytes);
void atmrt_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* ATMRT_H */
