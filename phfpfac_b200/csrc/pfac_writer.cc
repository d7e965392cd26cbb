// GPU_match_result.txt writer: one line per record,
//     "At position %4d, match pattern %d\n"            (reference main.cc:344)
// Positions are 64-bit here (the reference's int caps inputs at 2 GiB, main.cc:79); the
// width-4 right-justified rule of %4d is kept for every magnitude.
//
// Optional binary sidecar of the compact records (the reference writes text only): a 32-byte
// header {"PFACREC1", u32 version, u32 record bytes, u64 blocks, u64 records} followed by blocks
// {u64 base position, u64 count, count x pfac_match}, little-endian, in position order.  8 bytes per match
// instead of 35-50 bytes of text; the text file is a pure function of it (pfac_format_records).
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

constexpr char kSideMagic[8] = {'P', 'F', 'A', 'C', 'R', 'E', 'C', '1'};
constexpr uint32_t kSideVersion = 1;

struct SideHeader {
    char magic[8];
    uint32_t version, record_bytes;
    uint64_t n_blocks, n_records;
};
static_assert(sizeof(SideHeader) == 32, "sidecar header is 32 bytes");

struct Sidecar {
    FILE *f = nullptr;
    SideHeader h;
};

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

int pfac_sidecar_begin(const char *path, void **sidecar)
{
    if (!path || !sidecar) return pfac::set_error(PFAC_ERR_ARG, "bad arguments");
    FILE *f = fopen(path, "wb");
    if (!f) return pfac::set_error(PFAC_ERR_IO, "Open sidecar file failed: %s", path);
    Sidecar *s = new Sidecar;
    s->f = f;
    memcpy(s->h.magic, kSideMagic, 8);
    s->h.version = kSideVersion;
    s->h.record_bytes = (uint32_t)sizeof(pfac_match);
    s->h.n_blocks = s->h.n_records = 0;
    // the totals are patched in by pfac_sidecar_end: a file whose header still says 0 blocks was cut short
    if (fwrite(&s->h, sizeof s->h, 1, f) != 1) {
        fclose(f);
        delete s;
        return pfac::set_error(PFAC_ERR_IO, "write failed: %s", path);
    }
    *sidecar = s;
    return PFAC_OK;
}

int pfac_sidecar_records(void *sidecar, uint64_t base_pos, const pfac_match *records, uint64_t count)
{
    Sidecar *s = (Sidecar *)sidecar;
    if (!s || (!records && count)) return pfac::set_error(PFAC_ERR_ARG, "bad arguments");
    if (!count) return PFAC_OK;   // empty segments leave no block
    const uint64_t head[2] = {base_pos, count};
    if (fwrite(head, sizeof head, 1, s->f) != 1 || fwrite(records, sizeof(pfac_match), (size_t)count, s->f) != (size_t)count)
        return pfac::set_error(PFAC_ERR_IO, "write failed");
    s->h.n_blocks++;
    s->h.n_records += count;
    return PFAC_OK;
}

int pfac_sidecar_end(void *sidecar)
{
    Sidecar *s = (Sidecar *)sidecar;
    if (!s) return pfac::set_error(PFAC_ERR_ARG, "bad arguments");
    int rc = PFAC_OK;
    if (fseek(s->f, 0, SEEK_SET) != 0 || fwrite(&s->h, sizeof s->h, 1, s->f) != 1) rc = pfac::set_error(PFAC_ERR_IO, "write failed");
    if (fclose(s->f) != 0 && rc == PFAC_OK) rc = pfac::set_error(PFAC_ERR_IO, "close failed");
    delete s;
    return rc;
}

int pfac_sidecar_read(const char *path, uint64_t *pos, uint32_t *id, uint64_t cap, uint64_t *n_records)
{
    if (!path || !n_records) return pfac::set_error(PFAC_ERR_ARG, "bad arguments");
    FILE *f = fopen(path, "rb");
    if (!f) return pfac::set_error(PFAC_ERR_IO, "Open sidecar file failed: %s", path);
    struct Closer {
        FILE *f;
        ~Closer() { fclose(f); }
    } closer{f};
    SideHeader h;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, kSideMagic, 8) != 0 || h.version != kSideVersion ||
        h.record_bytes != sizeof(pfac_match))
        return pfac::set_error(PFAC_ERR_IO, "%s is not a record sidecar of this version", path);
    *n_records = h.n_records;
    if (!pos && !id) return PFAC_OK;   // size query
    if (cap < h.n_records) return pfac::set_error(PFAC_ERR_OUTPUT_FULL, "%llu records, capacity %llu", (unsigned long long)h.n_records, (unsigned long long)cap);
    std::vector<pfac_match> block;
    uint64_t done = 0;
    for (uint64_t b = 0; b < h.n_blocks; b++) {
        uint64_t head[2];
        if (fread(head, sizeof head, 1, f) != 1 || head[1] > h.n_records - done)
            return pfac::set_error(PFAC_ERR_IO, "%s: block %llu is damaged", path, (unsigned long long)b);
        for (uint64_t left = head[1]; left;) {   // bounded buffer: a block may hold 10^9 records
            const size_t n = (size_t)std::min<uint64_t>(left, 1u << 20);
            block.resize(n);
            if (fread(block.data(), sizeof(pfac_match), n, f) != n)
                return pfac::set_error(PFAC_ERR_IO, "%s: block %llu is cut short", path, (unsigned long long)b);
            for (size_t i = 0; i < n; i++) {
                if (pos) pos[done + i] = head[0] + block[i].pos;
                if (id) id[done + i] = block[i].id;
            }
            done += n;
            left -= n;
        }
    }
    if (done != h.n_records) return pfac::set_error(PFAC_ERR_IO, "%s: %llu records in the blocks, %llu in the header", path, (unsigned long long)done, (unsigned long long)h.n_records);
    return PFAC_OK;
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
