// atmrt_host.h -- host-side helpers of the thin C++ host (what the reference keeps on the CPU):
// DTED decoding (terrain/mod.rs:24,85-98 through the external crate dted 0.2), PNG output
// (renderer/mod.rs:433-436 through the `image` crate), YAML/CLI configuration
// (generator/params.rs) and the `gen` subcommand (generator/mod.rs:47-99).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../../include/atmrt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* read_dted_header / read_dted: bit-exact decode of a DTED file (MIL-PRF-89020B): big-endian
 * signed-magnitude posts, [lon line][lat point]. posts == NULL reads only the header. */
int atmrt_host_read_dted(const char* path, atmrt_tile_desc* desc, int16_t* posts, size_t capacity);
/* GeoTIFF terrain tiles (terrain/geotiff.rs; external crate geotiff-rs 0.1 for the pixels): the key and the south-west corner
 * come from the FILE NAME -- the first (N|S)(\d+)(E|W)(\d+) in it (geotiff.rs:16-31) -- and the tile becomes 3601 x 3601 posts
 * one arc-second apart, [lon line][lat point] west->east, south->north like a DTED tile, which DtedData::get_elev then samples
 * exactly as GeoTiffWrapper::get_elev does (geotiff.rs:62-99). posts == NULL reads only the descriptor. Classic TIFF, strips
 * or tiles, 8/16/32-bit integers, none / PackBits / LZW / Deflate, predictor 2; samples must fit i16. */
int atmrt_host_geotiff_coords_from_name(const char* path, int* lat, int* lon);
int atmrt_host_read_geotiff(const char* path, atmrt_tile_desc* desc, int16_t* posts, size_t capacity);
/* TerrainDataInner::read_tile / Terrain::buffer_file (terrain/mod.rs:23-31, 85-118): DTED first, GeoTIFF second. */
int atmrt_host_read_tile(const char* path, atmrt_tile_desc* desc, int16_t* posts, size_t capacity);
/* RGB8 or RGBA8 PNG writer/reader on zlib (channels = 3 or 4). */
int atmrt_host_write_png(const char* path, const uint8_t* pixels, int width, int height, int channels);
int atmrt_host_read_png(const char* path, uint8_t* rgba, size_t capacity, int* width, int* height);
const char* atmrt_host_last_error(void);
/* The writer of the metadata sidecar, on its own (tests): gzip as a series of members compressed on `threads` host threads
 * (0: all), block_bytes of input per member (0: 4 MiB). */
int atmrt_host_gzip_write(const char* path, const void* data, size_t bytes, size_t block_bytes, int threads);
/* The overlays renderer::output_image draws over the finished picture (renderer/mod.rs:22-365, 416-431): ticks with
 * labels, the flat-earth horizon line, the eye-level line. Tick / VerticalTick (params.rs:325-385) as one POD. */
typedef struct atmrt_host_tick {
    int32_t multiple;  /* 0: Single { azimuth | elevation, size, labelled }, 1: Multiple { bias, step, size, labelled } */
    int32_t labelled;
    uint32_t size;     /* pixels */
    uint32_t reserved;
    double angle;      /* Single: azimuth (ticks) / elevation (vertical_ticks), degrees */
    double bias, step; /* Multiple */
} atmrt_host_tick;
typedef struct atmrt_host_overlays {
    const atmrt_host_tick* ticks;          /* params.output.ticks */
    const atmrt_host_tick* vertical_ticks; /* params.output.vertical_ticks */
    int32_t nticks, nvertical_ticks;
    double direction, fov, tilt;           /* params.view.frame, degrees */
    int32_t show_eye_level;                /* params.output.show_eye_level */
    int32_t show_flat_horizon;             /* ALREADY gated as output_image gates it: show_flat_horizon && shape == Flat && !straight_rays */
    double flat_horizon_elevation;         /* degrees: atmrt_host_flat_horizon_elevation(n at the observer's altitude) */
} atmrt_host_overlays;
/* Draws into rgb[height][width][3] in output_image's order (ticks, flat horizon, eye level). elevation_angle / azimuth:
 * ResultPixel.elevation_angle / .azimuth as atmrt_pixel_angles returns them, [height][width] each. */
int atmrt_host_draw_overlays(uint8_t* rgb, int width, int height, const double* elevation_angle, const double* azimuth,
                             const atmrt_host_overlays* overlays);
/* The ticks draw_ticks draws (gen_ticks, renderer/mod.rs:225-266), by ascending pixel position: x for ticks (vertical == 0),
 * y for vertical_ticks. label = format!("{:.1$}", angle, decimals). n = how many there are (may exceed capacity). */
typedef struct atmrt_host_draw_tick {
    uint32_t position, size;
    int32_t labelled;
    char label[28];
} atmrt_host_draw_tick;
int atmrt_host_gen_ticks(int width, int height, const double* elevation_angle, const double* azimuth, const atmrt_host_overlays* overlays,
                         int vertical, atmrt_host_draw_tick* out, int capacity, int* n);
int atmrt_host_num_decimals(double x); /* renderer/mod.rs:206-214 (the reference's own test: renderer/mod.rs:438-459) */
int atmrt_host_flat_horizon_elevation(double n_at_observer, double* elevation_deg); /* renderer/mod.rs:424-425 */
/* The subcommands of the reference's binary on this path (main.rs:17-39), argv without the subcommand name; each
 * returns the process exit code. gen: generator/mod.rs:47-99; the three text dumpers: ray_path.rs, elev_profile.rs,
 * atm_printer.rs (same flags, same text layout; the numbers come from the device through the C-ABI probes). */
int atmrt_host_gen(int argc, const char* const* argv);
int atmrt_host_output_ray_paths(int argc, const char* const* argv);
int atmrt_host_output_elev_profile(int argc, const char* const* argv);
int atmrt_host_output_atm(int argc, const char* const* argv);
/* read_config + Config::into_params without touching the GPU (tests): YAML + CLI -> params / objects / paths. */
int atmrt_host_parse_config(int argc, const char* const* argv, atmrt_params* params, atmrt_object* objects, int max_objects,
                            int* nobjects, char* terrain_folder, size_t folder_cap, char* output_file, size_t file_cap,
                            char* meta_file, size_t meta_cap);

/* Parse-only: output.ticks / vertical_ticks / show_eye_level / show_flat_horizon as read_config lowers them (tests). */
int atmrt_host_parse_overlays(int argc, const char* const* argv, atmrt_host_tick* ticks, int max_ticks, int* nticks,
                              atmrt_host_tick* vertical_ticks, int max_vertical_ticks, int* nvertical_ticks, int* show_eye_level,
                              int* show_flat_horizon);

#ifdef __cplusplus
}
#endif
