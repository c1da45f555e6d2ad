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

#ifdef __cplusplus
}
#endif
