// pgzip.h -- a gzip file written as a series of members compressed on all host threads (RFC 1952 section 2.2: "a gzip file
// consists of a series of members"; zlib's gzread, gzip(1) and Python's gzip module read them back as one stream). The metadata
// sidecar of a c2 frame is 66 MB, of a c5 frame 2.1 GB: one deflate stream at level 6 runs at ~80 MB/s on one core.
#pragma once
#include <cstddef>
#include <cstdio>
#include <string>
#include <vector>

namespace atmrt_host {

class ParallelGzip {
public:
    explicit ParallelGzip(const std::string& path, int level = 6, size_t block_bytes = 4u << 20, unsigned threads = 0);
    ~ParallelGzip();
    ParallelGzip(const ParallelGzip&) = delete;
    ParallelGzip& operator=(const ParallelGzip&) = delete;
    bool ok() const { return ok_; }
    void put(const void* data, size_t bytes);
    bool close();  // compresses what is pending, closes the file; false if anything failed

private:
    void flush_wave();
    FILE* f_ = nullptr;
    bool ok_ = false, any_ = false;
    int level_;
    size_t block_;
    unsigned threads_;
    std::string cur_;
    std::vector<std::string> pending_;
};

}  // namespace atmrt_host
