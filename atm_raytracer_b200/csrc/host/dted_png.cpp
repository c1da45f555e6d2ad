// dted_png.cpp -- host file formats: DTED decode and PNG encode/decode (on zlib).
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "atmrt_host.h"

namespace atmrt_host {
std::string g_error;
int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
}  // namespace atmrt_host
using atmrt_host::fail;

namespace {

bool parse_uint(const unsigned char* s, int n, int* out) {
    int v = 0;
    for (int i = 0; i < n; ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (s[i] - '0');
    }
    *out = v;
    return true;
}

// "DDDMMSSH" -> signed decimal degrees
bool parse_angle(const unsigned char* s, double* out) {
    int d, m, sec;
    if (!parse_uint(s, 3, &d) || !parse_uint(s + 3, 2, &m) || !parse_uint(s + 5, 2, &sec)) return false;
    char h = (char)s[7];
    if (h != 'N' && h != 'S' && h != 'E' && h != 'W') return false;
    double v = (double)d + (double)m / 60.0 + (double)sec / 3600.0;
    *out = (h == 'S' || h == 'W') ? -v : v;
    return true;
}

int saturating_i16(double v) {  // Rust `as i16`
    if (!(v == v)) return 0;
    if (v <= -32768.0) return -32768;
    if (v >= 32767.0) return 32767;
    return (int)v;
}

}  // namespace

extern "C" int atmrt_host_read_dted(const char* path, atmrt_tile_desc* desc, int16_t* posts, size_t capacity) {
    if (!path || !desc) return fail(ATMRT_ERR_INVALID, "read_dted: NULL argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(ATMRT_ERR_IO, std::string("cannot open ") + path);
    unsigned char uhl[80];
    if (fread(uhl, 1, 80, f) != 80 || memcmp(uhl, "UHL1", 4) != 0) {
        fclose(f);
        return fail(ATMRT_ERR_INVALID, std::string(path) + ": not a DTED file (no UHL1 record)");
    }
    int lon_iv, lat_iv, nlon, nlat;
    double lon0, lat0;
    if (!parse_angle(uhl + 4, &lon0) || !parse_angle(uhl + 12, &lat0) || !parse_uint(uhl + 20, 4, &lon_iv) ||
        !parse_uint(uhl + 24, 4, &lat_iv) || !parse_uint(uhl + 47, 4, &nlon) || !parse_uint(uhl + 51, 4, &nlat) || nlon < 2 ||
        nlat < 2 || lon_iv <= 0 || lat_iv <= 0) {
        fclose(f);
        return fail(ATMRT_ERR_INVALID, std::string(path) + ": malformed UHL record");
    }
    desc->min_lon = lon0;
    desc->min_lat = lat0;
    desc->lon_interval = (double)lon_iv / 10.0;  // header unit: tenths of arc-seconds
    desc->lat_interval = (double)lat_iv / 10.0;
    desc->nlon = nlon;
    desc->nlat = nlat;
    desc->lat0 = saturating_i16(lat0);  // terrain/mod.rs:91-92
    desc->lon0 = saturating_i16(lon0);
    if (!posts) {
        fclose(f);
        return 0;
    }
    const size_t need = (size_t)nlon * (size_t)nlat;
    if (capacity < need) {
        fclose(f);
        return fail(ATMRT_ERR_INVALID, "read_dted: posts buffer too small");
    }
    const size_t rec = 12 + 2 * (size_t)nlat;  // 0xAA, 3 B block count, 2 B lon, 2 B lat, posts, 4 B checksum
    std::vector<unsigned char> buf(rec);
    if (fseek(f, 80 + 648 + 2700, SEEK_SET) != 0) {
        fclose(f);
        return fail(ATMRT_ERR_IO, std::string(path) + ": seek failed");
    }
    for (int i = 0; i < nlon; ++i) {
        if (fread(buf.data(), 1, rec, f) != rec || buf[0] != 0xAA) {
            fclose(f);
            return fail(ATMRT_ERR_INVALID, std::string(path) + ": truncated or corrupt data record " + std::to_string(i));
        }
        int16_t* out = posts + (size_t)i * nlat;
        for (int j = 0; j < nlat; ++j) {
            unsigned v = ((unsigned)buf[8 + 2 * j] << 8) | (unsigned)buf[9 + 2 * j];
            int mag = (int)(v & 0x7fffu);
            out[j] = (int16_t)((v & 0x8000u) ? -mag : mag);  // signed magnitude, not two's complement
        }
    }
    fclose(f);
    return 0;
}

// ---- PNG --------------------------------------------------------------------------------------
namespace {

void put_u32(std::vector<unsigned char>& v, uint32_t x) {
    v.push_back((unsigned char)(x >> 24));
    v.push_back((unsigned char)(x >> 16));
    v.push_back((unsigned char)(x >> 8));
    v.push_back((unsigned char)x);
}

void put_chunk(std::vector<unsigned char>& out, const char* type, const unsigned char* data, size_t n) {
    put_u32(out, (uint32_t)n);
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    uint32_t crc = (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4));
    put_u32(out, crc);
}

uint32_t get_u32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

// The image crate's encoder of the reference (renderer/mod.rs:433-436) is one deflate stream on one core; a 16384 x 4096 picture
// is 201 MB of scanlines -- seconds of deflate against 8 ms of rendering. Here the scanlines are deflated in bands on all host
// threads (raw deflate ended by a sync flush, so that the bands' outputs concatenate into ONE zlib stream -- the pigz
// construction --, Adler-32 of the whole by adler32_combine) and written as one IDAT chunk per band. Any PNG decoder reads it.
extern "C" int atmrt_host_write_png(const char* path, const uint8_t* pixels, int width, int height, int channels) {
    if (!path || !pixels || width <= 0 || height <= 0 || (channels != 3 && channels != 4))
        return fail(ATMRT_ERR_INVALID, "write_png: bad argument");
    const size_t stride = (size_t)width * channels;
    const int band_rows = (int)std::max<size_t>(1, std::min<size_t>((size_t)height, ((size_t)2 << 20) / (stride + 1) + 1));  // ~2 MB of scanlines
    const int nbands = (height + band_rows - 1) / band_rows;
    struct Band {
        std::vector<unsigned char> chunk;  // the whole IDAT chunk: length, "IDAT", data, CRC
        uLong adler = 0;
        size_t raw_len = 0;
        bool ok = false;
    };
    std::vector<Band> bands((size_t)nbands);
    std::atomic<int> next{0};
    auto work = [&] {
        std::vector<unsigned char> raw;
        for (int b; (b = next.fetch_add(1)) < nbands;) {
            Band& B = bands[(size_t)b];
            const int y0 = b * band_rows, y1 = std::min(height, y0 + band_rows);
            raw.resize((stride + 1) * (size_t)(y1 - y0));
            for (int y = y0; y < y1; ++y) {
                raw[(stride + 1) * (size_t)(y - y0)] = 0;  // filter type None
                memcpy(&raw[(stride + 1) * (size_t)(y - y0) + 1], pixels + stride * (size_t)y, stride);
            }
            B.raw_len = raw.size();
            B.adler = adler32(adler32(0L, Z_NULL, 0), raw.data(), (uInt)raw.size());
            z_stream z{};
            if (deflateInit2(&z, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) continue;
            const bool first = b == 0, last = b == nbands - 1;
            const size_t bound = deflateBound(&z, (uLong)raw.size()) + 16;
            B.chunk.resize(8 + (first ? 2 : 0) + bound + 4 + 4);
            size_t at = 8;
            if (first) B.chunk[at++] = 0x78, B.chunk[at++] = 0x9C;  // zlib header: deflate, 32 K window, default level
            z.next_in = raw.data(), z.avail_in = (uInt)raw.size();
            z.next_out = &B.chunk[at], z.avail_out = (uInt)bound;
            const int rc = deflate(&z, last ? Z_FINISH : Z_SYNC_FLUSH);
            const bool done = last ? rc == Z_STREAM_END : (rc == Z_OK && z.avail_in == 0 && z.avail_out > 0);
            at += z.total_out;
            deflateEnd(&z);
            if (!done) continue;
            B.chunk.resize(at + (last ? 4 : 0) + 4);  // (+ the stream's Adler-32, filled in below) + the chunk's CRC
            B.ok = true;
        }
    };
    {
        const unsigned nt = std::max(1u, std::min<unsigned>({std::thread::hardware_concurrency(), 64u, (unsigned)nbands}));
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    uLong adler = adler32(0L, Z_NULL, 0);
    for (const Band& B : bands) {
        if (!B.ok) return fail(ATMRT_ERR_IO, "write_png: deflate failed");
        adler = adler32_combine(adler, B.adler, (z_off_t)B.raw_len);
    }
    for (size_t b = 0; b < bands.size(); ++b) {  // chunk framing: length and type in front, (Adler-32 and) CRC behind
        std::vector<unsigned char>& c = bands[b].chunk;
        const bool last = b + 1 == bands.size();
        const size_t n = c.size() - 12;  // data bytes
        if (last) c[c.size() - 8] = (unsigned char)(adler >> 24), c[c.size() - 7] = (unsigned char)(adler >> 16), c[c.size() - 6] = (unsigned char)(adler >> 8), c[c.size() - 5] = (unsigned char)adler;
        c[0] = (unsigned char)(n >> 24), c[1] = (unsigned char)(n >> 16), c[2] = (unsigned char)(n >> 8), c[3] = (unsigned char)n;
        memcpy(&c[4], "IDAT", 4);
        uLong crc = crc32(0L, Z_NULL, 0);
        for (size_t o = 4; o < c.size() - 4;) {  // crc32 takes 32-bit lengths
            const size_t m = std::min<size_t>(c.size() - 4 - o, (size_t)1 << 30);
            crc = crc32(crc, &c[o], (uInt)m);
            o += m;
        }
        c[c.size() - 4] = (unsigned char)(crc >> 24), c[c.size() - 3] = (unsigned char)(crc >> 16), c[c.size() - 2] = (unsigned char)(crc >> 8), c[c.size() - 1] = (unsigned char)crc;
    }
    std::vector<unsigned char> head = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<unsigned char> ihdr;
    put_u32(ihdr, (uint32_t)width);
    put_u32(ihdr, (uint32_t)height);
    ihdr.push_back(8);
    ihdr.push_back(channels == 3 ? 2 : 6);
    ihdr.push_back(0);
    ihdr.push_back(0);
    ihdr.push_back(0);
    put_chunk(head, "IHDR", ihdr.data(), ihdr.size());
    std::vector<unsigned char> tail;
    put_chunk(tail, "IEND", nullptr, 0);
    FILE* f = fopen(path, "wb");
    if (!f) return fail(ATMRT_ERR_IO, std::string("cannot create ") + path);
    bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
    for (const Band& B : bands) ok = ok && fwrite(B.chunk.data(), 1, B.chunk.size(), f) == B.chunk.size();
    ok = ok && fwrite(tail.data(), 1, tail.size(), f) == tail.size();
    ok = fclose(f) == 0 && ok;
    return ok ? 0 : fail(ATMRT_ERR_IO, std::string("short write to ") + path);
}

// 8-bit, non-interlaced PNG (grey, grey+alpha, RGB, RGBA, palette) -> RGBA8.
extern "C" int atmrt_host_read_png(const char* path, uint8_t* rgba, size_t capacity, int* width, int* height) {
    if (!path || !width || !height) return fail(ATMRT_ERR_INVALID, "read_png: NULL argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(ATMRT_ERR_IO, std::string("cannot open ") + path);
    std::vector<unsigned char> file;
    unsigned char tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) file.insert(file.end(), tmp, tmp + n);
    fclose(f);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) return fail(ATMRT_ERR_INVALID, std::string(path) + ": not a PNG file");
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, plte, trns;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        uint32_t len = get_u32(&file[pos]);
        if (pos + 12 + (size_t)len > file.size()) return fail(ATMRT_ERR_INVALID, std::string(path) + ": truncated chunk");
        const unsigned char* type = &file[pos + 4];
        const unsigned char* data = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            w = (int)get_u32(data), h = (int)get_u32(data + 4);
            depth = data[8], ctype = data[9], interlace = data[12];
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!memcmp(type, "tRNS", 4)) {
            trns.assign(data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (w <= 0 || h <= 0) return fail(ATMRT_ERR_INVALID, std::string(path) + ": missing IHDR");
    if (depth != 8 || interlace != 0) return fail(ATMRT_ERR_INVALID, std::string(path) + ": only 8-bit non-interlaced PNGs are supported");
    int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) return fail(ATMRT_ERR_INVALID, std::string(path) + ": unsupported colour type");
    *width = w;
    *height = h;
    if (!rgba) return 0;
    if (capacity < (size_t)w * h * 4) return fail(ATMRT_ERR_INVALID, "read_png: buffer too small");
    const size_t stride = (size_t)w * ch;
    std::vector<unsigned char> raw((stride + 1) * (size_t)h);
    uLongf rlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rlen, idat.data(), (uLong)idat.size()) != Z_OK || rlen != raw.size())
        return fail(ATMRT_ERR_INVALID, std::string(path) + ": inflate failed");
    std::vector<unsigned char> img(stride * (size_t)h);
    for (int y = 0; y < h; ++y) {
        const unsigned char* in = &raw[(stride + 1) * y];
        unsigned char* out = &img[stride * y];
        const unsigned char* up = y ? &img[stride * (y - 1)] : nullptr;
        int ft = in[0];
        for (size_t i = 0; i < stride; ++i) {
            int a = i >= (size_t)ch ? out[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i - ch] : 0;
            int x = in[1 + i];
            switch (ft) {
                case 0: break;
                case 1: x += a; break;
                case 2: x += b; break;
                case 3: x += (a + b) / 2; break;
                case 4: x += paeth(a, b, c); break;
                default: return fail(ATMRT_ERR_INVALID, std::string(path) + ": bad filter type");
            }
            out[i] = (unsigned char)x;
        }
    }
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        const unsigned char* p = &img[i * ch];
        unsigned char r, g, b, a = 255;
        if (ctype == 0) {
            r = g = b = p[0];
        } else if (ctype == 4) {
            r = g = b = p[0], a = p[1];
        } else if (ctype == 2) {
            r = p[0], g = p[1], b = p[2];
        } else if (ctype == 6) {
            r = p[0], g = p[1], b = p[2], a = p[3];
        } else {
            size_t idx = p[0];
            if (idx * 3 + 2 >= plte.size()) return fail(ATMRT_ERR_INVALID, std::string(path) + ": palette index out of range");
            r = plte[idx * 3], g = plte[idx * 3 + 1], b = plte[idx * 3 + 2];
            a = idx < trns.size() ? trns[idx] : 255;
        }
        rgba[i * 4] = r, rgba[i * 4 + 1] = g, rgba[i * 4 + 2] = b, rgba[i * 4 + 3] = a;
    }
    return 0;
}

extern "C" const char* atmrt_host_last_error(void) { return atmrt_host::g_error.c_str(); }
