// Multi-GPU job: the GPU x stream scan loop of the reference (main.cc:171-272), re-cut for
// input sharding.  GPU g scans the contiguous chunk [start_g, end_g) plus a halo of
// max_pat_len-1 bytes and reports only matches that START inside its chunk, so concatenating
// the per-GPU record lists in GPU order is already globally position-ordered.  One host thread
// per GPU (the reference uses one OpenMP thread per GPU x stream, main.cc:225-241); streams
// inside a GPU are driven asynchronously by pfac_scan_host.  No inter-GPU traffic, no NCCL.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE   // O_DIRECT
#endif
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

#include "pfac_internal.h"

namespace {

struct Segment {
    uint64_t base = 0;       // global position of record pos 0
    uint64_t n_starts = 0;
    pfac_match *rec = nullptr;   // pinned
    uint64_t cap = 0;
    uint64_t count = 0;
};

// one pfac_scan_host call covers < 4 GiB; keep segments at 1 GiB of start positions
constexpr uint64_t kSegmentBytes = 1ull << 30;

}  // namespace

struct pfac_job {
    std::vector<pfac_ctx *> ctxs;
    std::vector<std::vector<Segment>> segs;   // per GPU
    std::vector<const Segment *> flat;
    int max_pat_len = 0;
    double secs[4] = {0, 0, 0, 0};
};

extern "C" {

int pfac_job_create(const pfac_tables *t, const int *devices, int n_devices, int streams_per_gpu,
                    size_t chunk_bytes, pfac_job **out)
{
    if (!t || !out || n_devices < 1 || streams_per_gpu < 1) return pfac::set_error(PFAC_ERR_ARG, "bad arguments to pfac_job_create");
    if (pfac_tables_n_parts(t) != 1)
        return pfac::set_error(PFAC_ERR_ARG, "the scanner takes a single-partition table set (got %d)", pfac_tables_n_parts(t));
    std::unique_ptr<pfac_job> job(new pfac_job);
    job->max_pat_len = pfac_tables_max_pat_len(t);
    for (int i = 0; i < n_devices; i++) {
        pfac_ctx *ctx = nullptr;
        int rc = pfac_ctx_create(devices ? devices[i] : i, t, 0, streams_per_gpu, chunk_bytes, &ctx);
        if (rc) {
            for (pfac_ctx *c : job->ctxs) pfac_ctx_destroy(c);
            return rc;
        }
        job->ctxs.push_back(ctx);
    }
    job->segs.resize((size_t)n_devices);
    *out = job.release();
    return PFAC_OK;
}

void pfac_job_destroy(pfac_job *job)
{
    if (!job) return;
    for (auto &v : job->segs)
        for (auto &s : v) pfac_host_free(s.rec);
    for (pfac_ctx *c : job->ctxs) pfac_ctx_destroy(c);
    delete job;
}

int pfac_job_plan(uint64_t n, int n_shards, int max_pat_len, int i, uint64_t *start, uint64_t *n_starts,
                  uint64_t *n_valid)
{
    if (n_shards < 1 || i < 0 || i >= n_shards || !start || !n_starts || !n_valid)
        return pfac::set_error(PFAC_ERR_ARG, "bad arguments to pfac_job_plan");
    const uint64_t halo = max_pat_len > 0 ? (uint64_t)max_pat_len - 1 : 0;
    // contiguous shard per GPU, 64 KiB granularity so sub-chunks stay tile aligned
    uint64_t per = (n + (uint64_t)n_shards - 1) / (uint64_t)n_shards;
    per = (per + 65535) & ~65535ull;
    const uint64_t lo = std::min<uint64_t>(n, (uint64_t)i * per), hi = std::min<uint64_t>(n, lo + per);
    *start = lo;
    *n_starts = hi - lo;
    *n_valid = std::min<uint64_t>(hi - lo + halo, n - lo);
    return PFAC_OK;
}

int pfac_job_run(pfac_job *job, const void *h_in, uint64_t n, uint64_t *n_matches)
{
    if (!job || (!h_in && n) || !n_matches) return pfac::set_error(PFAC_ERR_ARG, "bad arguments to pfac_job_run");
    const int G = (int)job->ctxs.size();
    const uint64_t halo = job->max_pat_len > 0 ? (uint64_t)job->max_pat_len - 1 : 0;
    std::vector<int> rcs((size_t)G, PFAC_OK);
    std::vector<std::string> errs((size_t)G);
    const auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int g) {
        uint64_t lo = 0, ns = 0, nvld = 0;
        pfac_job_plan(n, G, job->max_pat_len, g, &lo, &ns, &nvld);
        const uint64_t hi = lo + ns;
        std::vector<Segment> &segs = job->segs[(size_t)g];
        const size_t n_seg = (size_t)((hi - lo + kSegmentBytes - 1) / kSegmentBytes);
        for (size_t i = n_seg; i < segs.size(); i++) { pfac_host_free(segs[i].rec); }
        segs.resize(n_seg);
        for (size_t i = 0; i < n_seg; i++) {
            Segment &s = segs[i];
            s.base = lo + (uint64_t)i * kSegmentBytes;
            s.n_starts = std::min<uint64_t>(kSegmentBytes, hi - s.base);
            s.count = 0;
            const uint64_t nv = std::min<uint64_t>(s.n_starts + halo, n - s.base);
            for (int attempt = 0; attempt < 2; attempt++) {
                if (!s.rec) {
                    // room for one match per 256 input bytes to begin with (pinning memory is not free: a
                    // record per 8 bytes would pin as much again as the input); denser inputs report the
                    // size they need and the segment is scanned again
                    s.cap = std::max<uint64_t>(s.cap, std::max<uint64_t>(s.n_starts / 256, 65536));
                    int rc = pfac_host_alloc((void **)&s.rec, (size_t)s.cap * sizeof(pfac_match));
                    if (rc) { rcs[(size_t)g] = rc; errs[(size_t)g] = pfac_last_error(); return; }
                }
                int rc = pfac_scan_host(job->ctxs[(size_t)g], (const uint8_t *)h_in + s.base, s.n_starts, nv,
                                        s.base, s.rec, s.cap, &s.count);
                if (rc == PFAC_ERR_OUTPUT_FULL && attempt == 0) {   // dense matches: exact-size retry
                    pfac_host_free(s.rec);
                    s.rec = nullptr;
                    s.cap = s.count;
                    continue;
                }
                if (rc) { rcs[(size_t)g] = rc; errs[(size_t)g] = pfac_last_error(); return; }
                break;
            }
        }
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; g++) th.emplace_back(work, g);
    work(0);
    for (auto &t : th) t.join();
    job->secs[0] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    job->flat.clear();
    uint64_t total = 0;
    for (int g = 0; g < G; g++) {
        if (rcs[(size_t)g]) return pfac::set_error(rcs[(size_t)g], "GPU %d: %s", pfac_ctx_device(job->ctxs[(size_t)g]), errs[(size_t)g].c_str());
        for (const Segment &s : job->segs[(size_t)g]) {
            job->flat.push_back(&s);
            total += s.count;
        }
    }
    *n_matches = total;
    return PFAC_OK;
}

// The same job fed from a FILE instead of a host buffer that already holds it (the reference freads the
// whole file into pinned memory before the first byte is scanned, main.cc:147-155): per GPU one reader
// thread preads its shard chunk by chunk (with PFAC_READER_ODIRECT=1 through O_DIRECT, so that the page
// cache is not filled with a multi-GB input) into a ring of pinned buffers, while the GPU's host thread
// scans the chunks that are in (pfac_scan_host: H2D, kernels, D2H).  The scan starts with the first
// chunk; file size is bounded by the record buffers, not by host memory.
int pfac_job_run_file(pfac_job *job, const char *path, uint64_t n, uint64_t *n_matches)
{
    if (!job || !path || !n_matches) return pfac::set_error(PFAC_ERR_ARG, "bad arguments to pfac_job_run_file");
    const int G = (int)job->ctxs.size();
    const uint64_t halo = job->max_pat_len > 0 ? (uint64_t)job->max_pat_len - 1 : 0;
    constexpr uint64_t kChunk = 64ull << 20, kAlign = 4096;
    constexpr int kRing = 6, kReaders = 3;   // (one thread copies out of the page cache at ~6 GB/s; the PCIe link takes ~55)
    std::vector<int> rcs((size_t)G, PFAC_OK);
    std::vector<std::string> errs((size_t)G);
    const auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int g) {
        auto fail = [&](int rc, const std::string &why) { rcs[(size_t)g] = rc; errs[(size_t)g] = why; };
        uint64_t lo = 0, ns_all = 0, nv_all = 0;
        pfac_job_plan(n, G, job->max_pat_len, g, &lo, &ns_all, &nv_all);
        const uint64_t hi = lo + ns_all;
        std::vector<Segment> &segs = job->segs[(size_t)g];
        const size_t n_seg = (size_t)((hi - lo + kSegmentBytes - 1) / kSegmentBytes);
        for (size_t i = n_seg; i < segs.size(); i++) pfac_host_free(segs[i].rec);
        segs.resize(n_seg);
        if (hi == lo) return;
        // PFAC_READER_ODIRECT=1: bypass the page cache (inputs that are read once and are larger than host
        // memory is worth); default: buffered reads with sequential read-ahead, which also profit from a
        // file that is still in the cache
        const char *od = getenv("PFAC_READER_ODIRECT");
        bool direct0 = od && atoi(od) != 0;
        int fd0 = direct0 ? open(path, O_RDONLY | O_DIRECT) : -1;
        if (fd0 < 0) {
            direct0 = false;
            fd0 = open(path, O_RDONLY);
        }
        if (fd0 < 0) return fail(PFAC_ERR_IO, std::string("Open input file failed: ") + path);
        if (!direct0) posix_fadvise(fd0, (off_t)lo, (off_t)(hi - lo), POSIX_FADV_SEQUENTIAL);
        // ring of pinned chunk buffers
        const size_t buf_bytes = (size_t)(kChunk + halo + 2 * kAlign);
        uint8_t *bufs[kRing] = {};
        for (auto &b : bufs)
            if (pfac_host_alloc((void **)&b, buf_bytes)) {
                const std::string why = pfac_last_error();
                for (auto &x : bufs) pfac_host_free(x);   // (null entries are skipped)
                close(fd0);
                return fail(PFAC_ERR_NOMEM, why);
            }
        struct Slot { uint64_t off = 0, len = 0, skip = 0; bool full = false, err = false; } slots[kRing];
        std::mutex mu;
        std::condition_variable cv;
        const uint64_t n_chunks = (hi - lo + kChunk - 1) / kChunk;
        bool stop = false;
        auto read_chunks = [&](int r0) {
            int fd = r0 == 0 ? fd0 : open(path, O_RDONLY | (direct0 ? O_DIRECT : 0));
            bool direct = direct0;
            if (fd < 0) fd = open(path, O_RDONLY), direct = false;
            for (uint64_t c = (uint64_t)r0; c < n_chunks; c += kReaders) {
                Slot &sl = slots[c % kRing];
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return !sl.full || stop; });
                    if (stop) { if (fd >= 0) close(fd); return; }
                }
                const uint64_t off = lo + c * kChunk, want = std::min<uint64_t>(kChunk, hi - off);
                const uint64_t end = std::min<uint64_t>(off + want + halo, n);          // readable bytes incl. halo
                const uint64_t a_off = off & ~(kAlign - 1), a_end = (end + kAlign - 1) & ~(kAlign - 1);
                uint8_t *dst = bufs[c % kRing];
                uint64_t got = 0;
                bool err = false;
                while (a_off + got < end) {
                    ssize_t r = pread(fd, dst + got, (size_t)(a_end - a_off - got), (off_t)(a_off + got));
                    if (r < 0 && direct && errno == EINVAL) {   // the file system refuses O_DIRECT here: plain reads
                        close(fd);
                        fd = open(path, O_RDONLY);
                        direct = false;
                        if (fd < 0) { err = true; break; }
                        continue;
                    }
                    if (r <= 0) { err = a_off + got < end; break; }
                    got += (uint64_t)r;
                }
                std::lock_guard<std::mutex> lk(mu);
                sl.off = off;
                sl.len = want;
                sl.skip = off - a_off;
                sl.err = err;
                sl.full = true;
                cv.notify_all();
            }
            if (fd >= 0) close(fd);
        };
        std::vector<std::thread> readers;
        for (int r0 = 0; r0 < kReaders; r0++) readers.emplace_back(read_chunks, r0);
        for (size_t i = 0; i < n_seg; i++) {
            Segment &s = segs[i];
            s.base = lo + (uint64_t)i * kSegmentBytes;
            s.n_starts = std::min<uint64_t>(kSegmentBytes, hi - s.base);
            s.count = 0;
        }
        for (uint64_t c = 0; c < n_chunks && rcs[(size_t)g] == PFAC_OK; c++) {
            Slot &sl = slots[c % kRing];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return sl.full; });
            }
            if (sl.err) { fail(PFAC_ERR_IO, std::string("short read on ") + path); break; }
            Segment &s = segs[(size_t)((sl.off - lo) / kSegmentBytes)];
            const uint64_t nv = std::min<uint64_t>(sl.len + halo, n - sl.off);
            const uint8_t *src = bufs[c % kRing] + sl.skip;
            for (int attempt = 0; attempt < 3; attempt++) {
                if (!s.rec || s.cap < s.count + 4096) {   // (re)size the segment's record buffer, keeping what is in
                    const uint64_t cap2 = std::max<uint64_t>(std::max<uint64_t>(s.cap * 2, s.count + 65536), s.n_starts / 256);
                    pfac_match *nr = nullptr;
                    if (pfac_host_alloc((void **)&nr, (size_t)cap2 * sizeof(pfac_match))) { fail(PFAC_ERR_NOMEM, pfac_last_error()); break; }
                    if (s.rec) { memcpy(nr, s.rec, (size_t)s.count * sizeof(pfac_match)); pfac_host_free(s.rec); }
                    s.rec = nr;
                    s.cap = cap2;
                }
                uint64_t m = 0;
                int rc = pfac_scan_host(job->ctxs[(size_t)g], src, sl.len, nv, sl.off, s.rec + s.count, s.cap - s.count, &m);
                if (rc == PFAC_ERR_OUTPUT_FULL && attempt < 2) {   // dense matches: grow and scan the chunk again
                    s.cap = std::max<uint64_t>(s.cap, s.count + m);
                    pfac_match *nr = nullptr;
                    if (pfac_host_alloc((void **)&nr, (size_t)(s.count + m + 4096) * sizeof(pfac_match))) { fail(PFAC_ERR_NOMEM, pfac_last_error()); break; }
                    memcpy(nr, s.rec, (size_t)s.count * sizeof(pfac_match));
                    pfac_host_free(s.rec);
                    s.rec = nr;
                    s.cap = s.count + m + 4096;
                    continue;
                }
                if (rc) { fail(rc, pfac_last_error()); break; }
                const uint32_t bias = (uint32_t)(sl.off - s.base);   // positions of a segment are relative to its base
                if (bias)
                    for (uint64_t k = 0; k < m; k++) s.rec[s.count + k].pos += bias;
                s.count += m;
                break;
            }
            std::lock_guard<std::mutex> lk(mu);
            sl.full = false;
            cv.notify_all();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
            cv.notify_all();
        }
        for (auto &t : readers) t.join();
        for (auto &b : bufs) pfac_host_free(b);
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; g++) th.emplace_back(work, g);
    work(0);
    for (auto &t : th) t.join();
    job->secs[0] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    job->flat.clear();
    uint64_t total = 0;
    for (int g = 0; g < G; g++) {
        if (rcs[(size_t)g]) return pfac::set_error(rcs[(size_t)g], "GPU %d: %s", pfac_ctx_device(job->ctxs[(size_t)g]), errs[(size_t)g].c_str());
        for (const Segment &s : job->segs[(size_t)g]) {
            job->flat.push_back(&s);
            total += s.count;
        }
    }
    *n_matches = total;
    return PFAC_OK;
}

int pfac_job_n_segments(const pfac_job *job) { return job ? (int)job->flat.size() : 0; }

int pfac_job_segment(const pfac_job *job, int i, uint64_t *base_pos, const pfac_match **records, uint64_t *count)
{
    if (!job || i < 0 || i >= (int)job->flat.size() || !base_pos || !records || !count)
        return pfac::set_error(PFAC_ERR_ARG, "bad segment index");
    const Segment *s = job->flat[(size_t)i];
    *base_pos = s->base;
    *records = s->rec;
    *count = s->count;
    return PFAC_OK;
}

int pfac_job_last_timing(const pfac_job *job, double secs[4])
{
    if (!job || !secs) return pfac::set_error(PFAC_ERR_ARG, "bad arguments");
    for (int i = 0; i < 4; i++) secs[i] = job->secs[i];
    return PFAC_OK;
}

}  // extern "C"
