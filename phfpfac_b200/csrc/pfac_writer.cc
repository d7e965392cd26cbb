// GPU_match_result.txt writer: one line per record,
//     "At position %4d, match pattern %d\n"            (reference main.cc:344)
// Positions are 64-bit here (the reference's int caps inputs at 2 GiB, main.cc:79); the
// width-4 right-justified rule of %4d is kept for every magnitude.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "pfac_internal.h"

namespace {

struct Writer {
    FILE *f = nullptr;
    std::vector<char> buf;
    std::vector<std::vector<char>> tbuf;   // per-thread buffers of the parallel formatter
};

// records per block of the parallel formatter: each block is formatted by one thread, blocks are
// written in order, so the file is byte-identical to the sequential one
constexpr uint64_t kParBlock = 1u << 18;
constexpr uint64_t kParMin = 1u << 20;

inline char *put_u64(char *p, uint64_t v, int min_width)
{
    char tmp[24];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    for (int i = n; i < min_width; i++) *p++ = ' ';
    while (n) *p++ = tmp[--n];
    return p;
}

// worst case: 12 + 20 + 16 + 10 + 1 bytes
constexpr size_t kMaxLine = 64;

inline char *put_line(char *p, uint64_t pos, uint32_t id)
{
    memcpy(p, "At position ", 12);
    p = put_u64(p + 12, pos, 4);
    memcpy(p, ", match pattern ", 16);
    p += 16;
    if ((int32_t)id < 0) {   // %d of a negative int (never produced by the scanner)
        *p++ = '-';
        p = put_u64(p, (uint64_t)(-(int64_t)(int32_t)id), 1);
    } else {
        p = put_u64(p, id, 1);
    }
    *p++ = '\n';
    return p;
}

}  // namespace

extern "C" {

int pfac_write_begin(const char *path, void **writer)
{
    if (!path || !writer) return pfac::set_error(PFAC_ERR_ARG, "bad arguments");
    FILE *f = fopen(path, "w");   // main.cc:336
    if (!f) return pfac::set_error(PFAC_ERR_IO, "Open output file failed: %s", path);
    Writer *w = new Writer;
    w->f = f;
    w->buf.resize(4u << 20);
    *writer = w;
    return PFAC_OK;
}

int pfac_write_records(void *writer, uint64_t base_pos, const pfac_match *records, uint64_t count)
{
    Writer *w = (Writer *)writer;
    if (!w || (!records && count)) return pfac::set_error(PFAC_ERR_ARG, "bad arguments");
    if (count >= kParMin) {
        // many matches (10^6 .. 10^9): format blocks on all host threads, write them in order
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const unsigned nt = (unsigned)std::min<uint64_t>(hw, (count + kParBlock - 1) / kParBlock);
        w->tbuf.resize(nt);
        for (uint64_t base = 0; base < count; base += (uint64_t)nt * kParBlock) {
            std::vector<size_t> len(nt, 0);
            std::vector<std::thread> th;
            auto work = [&](unsigned t) {
                const uint64_t lo = base + (uint64_t)t * kParBlock;
                if (lo >= count) return;
                const uint64_t hi = std::min<uint64_t>(count, lo + kParBlock);
                std::vector<char> &b = w->tbuf[t];
                if (b.size() < (size_t)(kParBlock * kMaxLine)) b.resize((size_t)(kParBlock * kMaxLine));
                char *q = b.data();
                for (uint64_t i = lo; i < hi; i++) q = put_line(q, base_pos + records[i].pos, records[i].id);
                len[t] = (size_t)(q - b.data());
            };
            for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
            work(0);
            for (auto &x : th) x.join();
            for (unsigned t = 0; t < nt; t++)
                if (len[t] && fwrite(w->tbuf[t].data(), 1, len[t], w->f) != len[t])
                    return pfac::set_error(PFAC_ERR_IO, "write failed");
        }
        return PFAC_OK;
    }
    char *p = w->buf.data();
    char *const end = p + w->buf.size() - kMaxLine;
    for (uint64_t i = 0; i < count; i++) {
        p = put_line(p, base_pos + records[i].pos, records[i].id);
        if (p > end) {
            if (fwrite(w->buf.data(), 1, (size_t)(p - w->buf.data()), w->f) != (size_t)(p - w->buf.data()))
                return pfac::set_error(PFAC_ERR_IO, "write failed");
            p = w->buf.data();
        }
    }
    if (p != w->buf.data() && fwrite(w->buf.data(), 1, (size_t)(p - w->buf.data()), w->f) != (size_t)(p - w->buf.data()))
        return pfac::set_error(PFAC_ERR_IO, "write failed");
    return PFAC_OK;
}

int pfac_write_end(void *writer)
{
    Writer *w = (Writer *)writer;
    if (!w) return pfac::set_error(PFAC_ERR_ARG, "bad arguments");
    int rc = fclose(w->f) == 0 ? PFAC_OK : pfac::set_error(PFAC_ERR_IO, "close failed");
    delete w;
    return rc;
}

size_t pfac_format_records(uint64_t base_pos, const pfac_match *records, uint64_t count, char *buf, size_t buf_len)
{
    size_t need = 0;
    char line[kMaxLine];
    for (uint64_t i = 0; i < count; i++) {
        char *e = put_line(line, base_pos + records[i].pos, records[i].id);
        size_t n = (size_t)(e - line);
        if (buf && need + n <= buf_len) memcpy(buf + need, line, n);
        need += n;
    }
    return need;
}

}  // extern "C"
