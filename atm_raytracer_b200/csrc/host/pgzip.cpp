// pgzip.cpp -- see pgzip.h. Blocks of the input become independent gzip members (deflate with the gzip wrapper, windowBits 15 + 16),
// compressed a wave at a time by a few threads and written in order.
#include "pgzip.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <thread>

#include "atmrt_host.h"

namespace atmrt_host {
int fail(int code, const std::string& msg);

namespace {
bool gzip_member(const std::string& in, int level, std::string* out) {
    z_stream s{};
    if (deflateInit2(&s, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out->resize(deflateBound(&s, (uLong)in.size()) + 64);
    s.next_in = (Bytef*)in.data(), s.avail_in = (uInt)in.size();
    s.next_out = (Bytef*)&(*out)[0], s.avail_out = (uInt)out->size();
    const int rc = deflate(&s, Z_FINISH);
    const size_t n = s.total_out;
    deflateEnd(&s);
    if (rc != Z_STREAM_END) return false;
    out->resize(n);
    return true;
}
}  // namespace

ParallelGzip::ParallelGzip(const std::string& path, int level, size_t block_bytes, unsigned threads)
    : level_(level), block_(std::max<size_t>(block_bytes, 1u << 16)), threads_(threads ? threads : std::max(1u, std::thread::hardware_concurrency())) {
    block_ = std::min<size_t>(block_, 1u << 30);  // a member's input fits zlib's 32-bit counters
    threads_ = std::min(threads_, 64u);
    f_ = fopen(path.c_str(), "wb");
    ok_ = f_ != nullptr;
}

ParallelGzip::~ParallelGzip() {
    if (f_) fclose(f_);
}

void ParallelGzip::put(const void* data, size_t bytes) {
    const char* p = (const char*)data;
    while (ok_ && bytes > 0) {
        const size_t take = std::min(bytes, block_ - cur_.size());
        cur_.append(p, take);
        p += take, bytes -= take;
        if (cur_.size() == block_) {
            pending_.push_back(std::move(cur_));
            cur_.clear();
            if (pending_.size() >= 2 * (size_t)threads_) flush_wave();
        }
    }
}

void ParallelGzip::flush_wave() {
    if (!ok_ || pending_.empty()) {
        pending_.clear();
        return;
    }
    std::vector<std::string> out(pending_.size());
    std::atomic<size_t> next{0};
    std::atomic<bool> good{true};
    auto work = [&] {
        for (size_t i; (i = next.fetch_add(1)) < pending_.size();)
            if (!gzip_member(pending_[i], level_, &out[i])) good = false;
    };
    std::vector<std::thread> th;
    const size_t n = std::min<size_t>(threads_, pending_.size());
    for (size_t t = 1; t < n; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    ok_ = good;
    for (size_t i = 0; ok_ && i < out.size(); ++i) ok_ = fwrite(out[i].data(), 1, out[i].size(), f_) == out[i].size();
    any_ = any_ || !out.empty();
    pending_.clear();
}

bool ParallelGzip::close() {
    if (!cur_.empty() || (!any_ && pending_.empty())) pending_.push_back(std::move(cur_));  // an empty input is one empty member
    cur_.clear();
    flush_wave();
    if (f_) {
        ok_ = fclose(f_) == 0 && ok_;
        f_ = nullptr;
    }
    return ok_;
}

}  // namespace atmrt_host

// test hook: the bytes through the parallel writer (threads == 0: all host threads)
extern "C" int atmrt_host_gzip_write(const char* path, const void* data, size_t bytes, size_t block_bytes, int threads) {
    if (!path || (bytes > 0 && !data)) return atmrt_host::fail(ATMRT_ERR_INVALID, "gzip_write: NULL argument");
    atmrt_host::ParallelGzip z(path, 6, block_bytes ? block_bytes : (size_t)(4u << 20), threads > 0 ? (unsigned)threads : 0u);
    z.put(data, bytes);
    return z.close() ? 0 : atmrt_host::fail(ATMRT_ERR_IO, std::string("cannot write ") + path);
}
