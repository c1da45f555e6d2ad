// geotiff.cpp -- GeoTIFF terrain tiles (terrain/geotiff.rs:9-100) decoded on the host into the same tile
// descriptor + int16 posts a DTED tile becomes, so the device samples both with one code path.
//
// What the reference does with such a file: the key and the south-west corner come from the FILE NAME, first match of
// (N|S)(\d+)(E|W)(\d+) (geotiff.rs:16-31); the tile spans one degree (geotiff.rs:46-60); get_elev (geotiff.rs:62-99) is
//     lat' = (lat - min_lat) * 3600, lon' likewise; truncate; at 3600 step back one post and add 1 to the fraction;
//     e00 (1-fx)(1-fy) + e01 (1-fx) fy + e10 fx (1-fy) + e11 fx fy     with e = get_pixel(lon_int [+1], lat_int [+1]) as f64
// which is DtedData::get_elev on a tile of 3601 x 3601 posts one arc-second apart, operation for operation
// (device_math.cuh:tile_get_elev: the division by an interval of 1.0 is exact). So: nlon = nlat = 3601, intervals 1.0.
//
// PARITY UNPINNED on one point: the pixels come from the external crate geotiff-rs 0.1 (Cargo.toml:13, not vendored with the
// reference). Its get_pixel(lon, lat) is taken in the sense the wrapper's arithmetic gives it -- the post `lat` arc-seconds
// NORTH of the tile's south edge and `lon` arc-seconds east of its west edge. A GeoTIFF raster starts at its north-west
// corner (ModelTiepoint), so post (lon, lat) is raster column lon of raster row 3600 - lat.
//
// The TIFF container is read per the TIFF 6.0 specification: classic (not Big) TIFF, either byte order, strips or tiles,
// one sample per pixel of 8 / 16 / 32 bits, signed or unsigned integers, compression none / PackBits / LZW / Deflate,
// horizontal-differencing predictor. Values outside i16 (the post type of the device terrain) and floating-point rasters are
// refused with a message.
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "atmrt_host.h"

namespace atmrt_host {
int fail(int code, const std::string& msg);
}
using atmrt_host::fail;

namespace {

constexpr int GEOTIFF_POSTS = 3601;  // geotiff.rs:71-88: indices 0 ..= 3600 along both axes

// GeoTiffWrapper::coords_from_name: the leftmost match of (N|S)(\d+)(E|W)(\d+) in the file name; a number i16::from_str
// refuses makes the whole name "not a GeoTIFF tile" (the `?` on .ok()), it does not look for a later match
bool coords_from_name(const std::string& path, int* lat, int* lon) {
    const size_t slash = path.find_last_of('/');
    const std::string name = slash == std::string::npos ? path : path.substr(slash + 1);
    auto digits = [&](size_t i) {
        size_t j = i;
        while (j < name.size() && name[j] >= '0' && name[j] <= '9') ++j;
        return j;
    };
    auto to_i16 = [&](size_t a, size_t b, int* out) {  // i16::from_str of ASCII digits: no sign here, overflow is an error
        long v = 0;
        for (size_t i = a; i < b; ++i) {
            v = v * 10 + (name[i] - '0');
            if (v > 32767) return false;
        }
        *out = (int)v;
        return true;
    };
    for (size_t i = 0; i < name.size(); ++i) {
        if (name[i] != 'N' && name[i] != 'S') continue;
        const size_t d1 = digits(i + 1);
        if (d1 == i + 1 || d1 >= name.size() || (name[d1] != 'E' && name[d1] != 'W')) continue;
        const size_t d2 = digits(d1 + 1);
        if (d2 == d1 + 1) continue;
        int la, lo;
        if (!to_i16(i + 1, d1, &la) || !to_i16(d1 + 1, d2, &lo)) return false;
        *lat = name[i] == 'S' ? -la : la;
        *lon = name[d1] == 'W' ? -lo : lo;
        return true;
    }
    return false;
}

struct Reader {
    std::vector<unsigned char> buf;
    bool big = false;
    uint32_t u16(size_t o) const {
        if (o + 2 > buf.size()) throw std::runtime_error("truncated file");
        return big ? (uint32_t)buf[o] << 8 | buf[o + 1] : (uint32_t)buf[o + 1] << 8 | buf[o];
    }
    uint32_t u32(size_t o) const {
        if (o + 4 > buf.size()) throw std::runtime_error("truncated file");
        return big ? (uint32_t)buf[o] << 24 | (uint32_t)buf[o + 1] << 16 | (uint32_t)buf[o + 2] << 8 | buf[o + 3]
                   : (uint32_t)buf[o + 3] << 24 | (uint32_t)buf[o + 2] << 16 | (uint32_t)buf[o + 1] << 8 | buf[o];
    }
};

struct Entry {
    uint32_t type = 0, count = 0;
    size_t value_at = 0;  // where the values are (inline in the entry, or at the offset it holds)
};

// values of a BYTE / SHORT / LONG entry
std::vector<uint32_t> values(const Reader& r, const Entry& e) {
    std::vector<uint32_t> out(e.count);
    for (uint32_t i = 0; i < e.count; ++i) {
        if (e.type == 1) {
            if (e.value_at + i >= r.buf.size()) throw std::runtime_error("truncated file");
            out[i] = r.buf[e.value_at + i];
        } else if (e.type == 3)
            out[i] = r.u16(e.value_at + 2 * (size_t)i);
        else if (e.type == 4)
            out[i] = r.u32(e.value_at + 4 * (size_t)i);
        else
            throw std::runtime_error("unexpected field type " + std::to_string(e.type));
    }
    return out;
}

// PackBits (TIFF 6.0 section 9)
std::vector<unsigned char> unpackbits(const unsigned char* p, size_t n, size_t want) {
    std::vector<unsigned char> out;
    out.reserve(want);
    size_t i = 0;
    while (i < n && out.size() < want) {
        const int c = (signed char)p[i++];
        if (c >= 0) {
            if (i + (size_t)c + 1 > n) throw std::runtime_error("PackBits: truncated run");
            out.insert(out.end(), p + i, p + i + c + 1);
            i += (size_t)c + 1;
        } else if (c != -128) {
            if (i >= n) throw std::runtime_error("PackBits: truncated run");
            out.insert(out.end(), (size_t)(1 - c), p[i++]);
        }
    }
    return out;
}

// LZW as TIFF uses it (TIFF 6.0 section 13): codes most significant bit first, 9 to 12 bits, ClearCode 256, EndOfInformation 257,
// the code width grows one code early ("early change")
std::vector<unsigned char> unlzw(const unsigned char* p, size_t n, size_t want) {
    std::vector<unsigned char> out;
    out.reserve(want);
    std::vector<uint32_t> prefix(4096);  // table entry: prefix code + last byte; length kept to write strings back to front
    std::vector<unsigned char> last(4096);
    std::vector<uint32_t> length(4096);
    for (int i = 0; i < 256; ++i) prefix[(size_t)i] = 0xffffffffu, last[(size_t)i] = (unsigned char)i, length[(size_t)i] = 1;
    uint32_t next = 258, width = 9, old = 0xffffffffu;
    uint64_t bits = 0;
    int nbits = 0;
    size_t i = 0;
    auto emit = [&](uint32_t code) {
        const size_t at = out.size(), len = length[code];
        out.resize(at + len);
        for (size_t k = len; k-- > 0; code = prefix[code]) out[at + k] = last[code];
    };
    auto first_byte = [&](uint32_t code) {
        while (prefix[code] != 0xffffffffu) code = prefix[code];
        return last[code];
    };
    while (out.size() < want) {
        while (nbits < (int)width && i < n) bits = bits << 8 | p[i++], nbits += 8;
        if (nbits < (int)width) break;  // ran out of data without an EndOfInformation code
        const uint32_t code = (uint32_t)(bits >> (nbits - (int)width)) & ((1u << width) - 1u);
        nbits -= (int)width;
        if (code == 257) break;
        if (code == 256) {
            next = 258, width = 9, old = 0xffffffffu;
            continue;
        }
        if (old == 0xffffffffu) {
            if (code >= 256) throw std::runtime_error("LZW: the first code after a clear is not a literal");
            emit(code);
        } else {
            if (code < next)
                emit(code);
            else if (code == next) {  // the string being defined: old + first byte of old
                emit(old);
                out.push_back(first_byte(old));
            } else
                throw std::runtime_error("LZW: code beyond the table");
            if (next < 4096) {
                prefix[next] = old, last[next] = first_byte(code < next ? code : old), length[next] = length[old] + 1;
                ++next;
            }
        }
        old = code;
        if (next + 1 >= (1u << width) && width < 12) ++width;  // early change: 511 -> 10 bits, 1023 -> 11, 2047 -> 12
    }
    return out;
}

std::vector<unsigned char> inflate_all(const unsigned char* p, size_t n, size_t want) {
    std::vector<unsigned char> out(want);
    uLongf len = (uLongf)want;
    const int rc = uncompress(out.data(), &len, p, (uLong)n);
    if (rc != Z_OK && rc != Z_BUF_ERROR) throw std::runtime_error("Deflate: zlib error " + std::to_string(rc));
    out.resize(len);
    return out;
}

struct Raster {
    uint32_t width = 0, height = 0;
    std::vector<int32_t> px;  // [row][column], row 0 = the top of the picture
};

Raster read_tiff(const std::string& path, bool header_only) {
    Reader r;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open the file");
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    r.buf.resize((size_t)std::max(size, 0L));
    if (size <= 0 || fread(r.buf.data(), 1, r.buf.size(), f) != r.buf.size()) {
        fclose(f);
        throw std::runtime_error("cannot read the file");
    }
    fclose(f);
    if (r.buf.size() < 8 || !((r.buf[0] == 'I' && r.buf[1] == 'I') || (r.buf[0] == 'M' && r.buf[1] == 'M'))) throw std::runtime_error("not a TIFF file");
    r.big = r.buf[0] == 'M';
    const uint32_t magic = r.u16(2);
    if (magic == 43) throw std::runtime_error("BigTIFF is not supported");
    if (magic != 42) throw std::runtime_error("not a TIFF file");
    const size_t ifd = r.u32(4);
    const uint32_t nent = r.u16(ifd);
    std::map<uint32_t, Entry> tags;
    static const size_t type_size[] = {0, 1, 1, 2, 4, 8, 1, 1, 2, 4, 8, 4, 8};
    for (uint32_t i = 0; i < nent; ++i) {
        const size_t at = ifd + 2 + 12 * (size_t)i;
        Entry e;
        const uint32_t tag = r.u16(at);
        e.type = r.u16(at + 2), e.count = r.u32(at + 4);
        const size_t bytes = (e.type < 13 ? type_size[e.type] : 0) * (size_t)e.count;
        e.value_at = bytes <= 4 ? at + 8 : (size_t)r.u32(at + 8);
        tags[tag] = e;
    }
    auto one = [&](uint32_t tag, uint32_t def, bool required = false) {
        auto it = tags.find(tag);
        if (it == tags.end()) {
            if (required) throw std::runtime_error("TIFF field " + std::to_string(tag) + " is missing");
            return def;
        }
        const std::vector<uint32_t> v = values(r, it->second);
        if (v.empty()) throw std::runtime_error("TIFF field " + std::to_string(tag) + " is empty");
        return v[0];
    };
    Raster img;
    img.width = one(256, 0, true), img.height = one(257, 0, true);
    const uint32_t bits = one(258, 1), compression = one(259, 1), spp = one(277, 1), predictor = one(317, 1), format = one(339, 1);
    if (spp != 1) throw std::runtime_error("expected one sample per pixel, the file has " + std::to_string(spp));
    if (format == 3) throw std::runtime_error("floating-point rasters are not supported (the device terrain holds i16 posts)");
    if (format != 1 && format != 2) throw std::runtime_error("unsupported SampleFormat " + std::to_string(format));
    if (bits != 8 && bits != 16 && bits != 32) throw std::runtime_error("unsupported BitsPerSample " + std::to_string(bits));
    if (predictor != 1 && predictor != 2) throw std::runtime_error("unsupported Predictor " + std::to_string(predictor));
    if (img.width == 0 || img.height == 0 || img.width > 65536 || img.height > 65536) throw std::runtime_error("unreasonable image size");
    if (header_only) return img;

    // chunks: strips (full-width bands of rows_per_strip rows) or tiles
    const bool tiled = tags.count(322) != 0;
    uint32_t cw, ch;
    std::vector<uint32_t> offsets, counts;
    if (tiled) {
        cw = one(322, 0, true), ch = one(323, 0, true);
        offsets = values(r, tags.at(324));
        if (!tags.count(325)) throw std::runtime_error("TileByteCounts is missing");
        counts = values(r, tags.at(325));
    } else {
        cw = img.width, ch = std::min(one(278, 0xffffffffu), img.height);
        if (!tags.count(273) || !tags.count(279)) throw std::runtime_error("StripOffsets / StripByteCounts is missing");
        offsets = values(r, tags.at(273)), counts = values(r, tags.at(279));
    }
    if (cw == 0 || ch == 0) throw std::runtime_error("empty strips / tiles");
    const uint32_t across = (img.width + cw - 1) / cw, down = (img.height + ch - 1) / ch;
    if (offsets.size() < (size_t)across * down || counts.size() < (size_t)across * down) throw std::runtime_error("too few strips / tiles");
    const size_t bps = bits / 8;
    img.px.assign((size_t)img.width * img.height, 0);
    for (uint32_t cy = 0; cy < down; ++cy)
        for (uint32_t cx = 0; cx < across; ++cx) {
            const size_t k = (size_t)cy * across + cx;
            const uint32_t rows = tiled ? ch : std::min(ch, img.height - cy * ch);  // a tile is always whole, the last strip may be short
            const size_t want = (size_t)cw * rows * bps;
            if ((size_t)offsets[k] + counts[k] > r.buf.size()) throw std::runtime_error("a strip / tile lies outside the file");
            const unsigned char* src = r.buf.data() + offsets[k];
            std::vector<unsigned char> raw;
            switch (compression) {
                case 1: raw.assign(src, src + counts[k]); break;
                case 5: raw = unlzw(src, counts[k], want); break;
                case 8:
                case 32946: raw = inflate_all(src, counts[k], want); break;
                case 32773: raw = unpackbits(src, counts[k], want); break;
                default: throw std::runtime_error("unsupported Compression " + std::to_string(compression));
            }
            if (raw.size() < want) throw std::runtime_error("a strip / tile decodes to fewer bytes than it holds pixels");
            for (uint32_t y = 0; y < rows; ++y) {
                const uint32_t iy = cy * ch + y;
                if (iy >= img.height) break;
                uint32_t acc = 0;  // horizontal differencing: sample = previous sample of the row + stored value, modulo 2^bits
                for (uint32_t x = 0; x < cw; ++x) {
                    const unsigned char* s = raw.data() + ((size_t)y * cw + x) * bps;
                    uint32_t v = bps == 1 ? s[0]
                                 : bps == 2 ? (r.big ? (uint32_t)s[0] << 8 | s[1] : (uint32_t)s[1] << 8 | s[0])
                                            : (r.big ? (uint32_t)s[0] << 24 | (uint32_t)s[1] << 16 | (uint32_t)s[2] << 8 | s[3]
                                                     : (uint32_t)s[3] << 24 | (uint32_t)s[2] << 16 | (uint32_t)s[1] << 8 | s[0]);
                    if (predictor == 2) v = acc = (acc + v) & (bits == 32 ? 0xffffffffu : (1u << bits) - 1u);
                    const uint32_t ix = cx * cw + x;
                    if (ix >= img.width) continue;
                    int64_t sv = v;
                    if (format == 2) sv = bits == 8 ? (int8_t)v : bits == 16 ? (int16_t)v : (int32_t)v;
                    if (sv < -32768 || sv > 32767) throw std::runtime_error("a sample (" + std::to_string(sv) + ") does not fit the i16 posts of the device terrain");
                    img.px[(size_t)iy * img.width + ix] = (int32_t)sv;
                }
            }
        }
    return img;
}

}  // namespace

extern "C" int atmrt_host_geotiff_coords_from_name(const char* path, int* lat, int* lon) {
    if (!path || !lat || !lon) return fail(ATMRT_ERR_INVALID, "geotiff_coords_from_name: NULL argument");
    if (!coords_from_name(path, lat, lon)) return fail(ATMRT_ERR_INVALID, std::string(path) + ": the file name holds no (N|S)<deg>(E|W)<deg>");
    return 0;
}

extern "C" int atmrt_host_read_geotiff(const char* path, atmrt_tile_desc* desc, int16_t* posts, size_t capacity) {
    if (!path || !desc) return fail(ATMRT_ERR_INVALID, "read_geotiff: NULL argument");
    int lat, lon;
    if (!coords_from_name(path, &lat, &lon)) return fail(ATMRT_ERR_INVALID, std::string(path) + ": the file name holds no (N|S)<deg>(E|W)<deg>");
    try {
        const Raster img = read_tiff(path, posts == nullptr);
        if (img.width < (uint32_t)GEOTIFF_POSTS || img.height < (uint32_t)GEOTIFF_POSTS)
            return fail(ATMRT_ERR_INVALID, std::string(path) + ": " + std::to_string(img.width) + " x " + std::to_string(img.height) +
                                               " pixels; the reference addresses 3601 x 3601 one-arc-second posts (geotiff.rs:66-88)");
        desc->lat0 = lat, desc->lon0 = lon;  // the HashMap key (terrain/mod.rs:100-111)
        desc->nlon = desc->nlat = GEOTIFF_POSTS;
        desc->min_lat = (double)lat, desc->min_lon = (double)lon;  // geotiff.rs:36-39; max = min + 1.0 (geotiff.rs:50-60)
        desc->lat_interval = desc->lon_interval = 1.0;             // arc-seconds: the 3600.0 of geotiff.rs:69-70
        if (!posts) return 0;
        const size_t need = (size_t)GEOTIFF_POSTS * GEOTIFF_POSTS;
        if (capacity < need) return fail(ATMRT_ERR_INVALID, "read_geotiff: posts buffer too small");
        // posts[lon line][lat point], south -> north: raster row 3600 - lat (see the note at the top of this file)
        const int B = 64;  // a transpose: in blocks, so that neither side walks the whole array with a stride
        for (int lo0 = 0; lo0 < GEOTIFF_POSTS; lo0 += B)
            for (int la0 = 0; la0 < GEOTIFF_POSTS; la0 += B)
                for (int lo = lo0; lo < std::min(lo0 + B, GEOTIFF_POSTS); ++lo)
                    for (int la = la0; la < std::min(la0 + B, GEOTIFF_POSTS); ++la)
                        posts[(size_t)lo * GEOTIFF_POSTS + la] = (int16_t)img.px[(size_t)(GEOTIFF_POSTS - 1 - la) * img.width + lo];
        return 0;
    } catch (const std::exception& e) {
        return fail(ATMRT_ERR_INVALID, std::string(path) + ": " + e.what());
    }
}

// Terrain::buffer_file / TerrainDataInner::read_tile (terrain/mod.rs:23-31, 85-118): a file is a DTED tile if its header reads as
// one, else a GeoTIFF tile if its name carries the coordinates; anything else stops the program ("Could not buffer terrain file").
extern "C" int atmrt_host_read_tile(const char* path, atmrt_tile_desc* desc, int16_t* posts, size_t capacity) {
    if (!path || !desc) return fail(ATMRT_ERR_INVALID, "read_tile: NULL argument");
    atmrt_tile_desc header{};
    if (atmrt_host_read_dted(path, &header, nullptr, 0) == 0) return atmrt_host_read_dted(path, desc, posts, capacity);
    int lat, lon;
    if (coords_from_name(path, &lat, &lon)) return atmrt_host_read_geotiff(path, desc, posts, capacity);
    return fail(ATMRT_ERR_INVALID, std::string("Could not buffer terrain file ") + path);
}
