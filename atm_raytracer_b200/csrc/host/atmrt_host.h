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
/* RGB8 or RGBA8 PNG writer/reader on zlib (channels = 3 or 4). */
int atmrt_host_write_png(const char* path, const uint8_t* pixels, int width, int height, int channels);
int atmrt_host_read_png(const char* path, uint8_t* rgba, size_t capacity, int* width, int* height);
const char* atmrt_host_last_error(void);
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

#ifdef __cplusplus
}
#endif
