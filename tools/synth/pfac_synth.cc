// Seeded synthetic workloads of BASELINE.json's configs.
// Input generation only -- no matching logic lives here.  Deterministic for a given
// (kind, seed, size): text is produced in independent 64 KiB blocks seeded by
// splitmix64(seed, block index), so the result does not depend on the thread count.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "pfac_synth.h"

namespace {

struct Rng {   // xoshiro256** seeded through splitmix64
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x)
    {
        uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed, uint64_t stream = 0)
    {
        uint64_t x = seed * 0xD1342543DE82EF95ULL + stream * 0x2545F4914F6CDD1DULL + 0x1234567ULL;
        for (auto &v : s) v = splitmix(x);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next()
    {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    int range(int lo, int hi) { return lo + (int)below((uint32_t)(hi - lo + 1)); }   // inclusive
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// token vocabulary shared by the Snort-like pattern set and the HTTP-like text (config 3)
const char *const kPrefixTok[] = {
    "GET /", "POST /", "HEAD /", "PUT /", "User-Agent: ", "Host: ", "Cookie: ", "Referer: ",
    "Content-Type: ", "Content-Length: ", "Accept: ", "Accept-Encoding: ", "Authorization: Basic ",
    "Connection: ", "X-Forwarded-For: ", "/cgi-bin/", "/admin/", "/wp-content/", "/wp-admin/",
    "/scripts/", "/phpmyadmin/", "/etc/passwd", "/bin/sh", "cmd.exe", ".php?", ".asp?", ".jsp?",
    ".cgi?", "id=", "cmd=", "file=", "page=", "SELECT ", "UNION ", "INSERT INTO ", "DROP TABLE ",
    "<script>", "javascript:", "onerror=", "../", "..\\", "%00", "%2e%2e/", "Mozilla/5.0 ",
    "curl/", "Wget/", "sqlmap/", "nikto", "HTTP/1.1", "HTTP/1.0", "application/", "text/html",
    "multipart/form-data", "boundary=", "charset=", "keep-alive", "gzip, deflate", "passwd=",
    "login=", "token=", "session=", "PHPSESSID=", "JSESSIONID=", "base64,",
};
const int kNumPrefixTok = (int)(sizeof(kPrefixTok) / sizeof(kPrefixTok[0]));
const char kAlnum[] = "abcdefghijklmnopqrstuvwxyz0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ_-./=&%";
const int kNumAlnum = (int)sizeof(kAlnum) - 1;

int lognormal_len(Rng &g, double median, double sigma, int lo, int hi)
{
    double u1 = g.unit(), u2 = g.unit();
    if (u1 < 1e-300) u1 = 1e-300;
    double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    double v = median * std::exp(sigma * z);
    int L = (int)std::lround(v);
    return std::min(hi, std::max(lo, L));
}

void append_http_line(Rng &g, std::string &out)
{
    switch (g.below(6)) {
    case 0: {
        out += kPrefixTok[g.below(4)];
        int segs = g.range(1, 3);
        for (int s = 0; s < segs; s++) {
            int L = g.range(3, 10);
            for (int i = 0; i < L; i++) out += kAlnum[g.below(36)];
            out += (s + 1 < segs) ? "/" : "";
        }
        if (g.below(2)) {
            out += kPrefixTok[24 + g.below(4)];
            out += kPrefixTok[28 + g.below(4)];
            int L = g.range(1, 8);
            for (int i = 0; i < L; i++) out += kAlnum[26 + g.below(10)];
        }
        out += " HTTP/1.1\r\n";
        break;
    }
    case 1:
        out += "Host: www.";
        for (int i = 0, L = g.range(4, 12); i < L; i++) out += kAlnum[g.below(26)];
        out += g.below(2) ? ".com\r\n" : ".org\r\n";
        break;
    case 2:
        out += "User-Agent: Mozilla/5.0 (X11; Linux x86_64) AppleWebKit/";
        for (int i = 0; i < 3; i++) out += kAlnum[26 + g.below(10)];
        out += ".36 (KHTML, like Gecko)\r\n";
        break;
    case 3:
        out += "Cookie: PHPSESSID=";
        for (int i = 0, L = g.range(16, 32); i < L; i++) out += kAlnum[g.below(36)];
        out += "; token=";
        for (int i = 0, L = g.range(8, 16); i < L; i++) out += kAlnum[g.below(62)];
        out += "\r\n";
        break;
    case 4:
        out += "Accept-Encoding: gzip, deflate\r\nConnection: keep-alive\r\nContent-Type: ";
        out += g.below(2) ? "application/x-www-form-urlencoded\r\n" : "text/html; charset=utf-8\r\n";
        break;
    default:
        out += "Content-Length: ";
        for (int i = 0, L = g.range(1, 5); i < L; i++) out += kAlnum[26 + g.below(10)];
        out += "\r\n\r\n";
        for (int i = 0, L = g.range(8, 48); i < L; i++) out += kAlnum[g.below((uint32_t)kNumAlnum)];
        out += "\r\n";
        break;
    }
}

constexpr size_t kBlock = 65536;   // one plant per block (every 64 KiB)

struct PatternView {
    std::vector<std::pair<const uint8_t *, int>> pats;
    explicit PatternView(const uint8_t *buf, size_t len)
    {
        size_t i = 0;
        while (i < len) {
            const uint8_t *nl = (const uint8_t *)memchr(buf + i, '\n', len - i);
            if (!nl) break;
            if (nl > buf + i) pats.push_back({buf + i, (int)(nl - (buf + i))});
            i = (size_t)(nl - buf) + 1;
        }
    }
};

void fill_block(int kind, uint64_t seed, uint64_t blk, uint8_t *out, size_t n, const PatternView *pv)
{
    Rng g(seed, blk + 1);
    if (kind == PFAC_SYNTH_TEXT_PRINTABLE) {
        // bytes uniform in 0x20..0x7E with '\n' every <= 120 bytes
        size_t i = 0;
        while (i < n) {
            size_t line = (size_t)g.range(40, 120);
            for (size_t k = 0; k + 1 < line && i < n; k++) out[i++] = (uint8_t)(0x20 + g.below(95));
            if (i < n) out[i++] = '\n';
        }
    } else {
        // HTTP-like lines from the vocabulary mixed 50/50 (by bytes) with uniform bytes
        size_t i = 0;
        std::string line;
        while (i < n) {
            size_t seg = (size_t)g.range(256, 1024);
            if (seg > n - i) seg = n - i;
            if (g.below(2)) {
                size_t end = i + seg;
                while (i < end) {
                    line.clear();
                    append_http_line(g, line);
                    size_t c = std::min(line.size(), end - i);
                    memcpy(out + i, line.data(), c);
                    i += c;
                }
            } else {
                size_t end = i + seg;
                while (i + 8 <= end) { uint64_t v = g.next(); memcpy(out + i, &v, 8); i += 8; }
                while (i < end) out[i++] = (uint8_t)g.below(256);
            }
        }
    }
    if (pv && !pv->pats.empty()) {   // planted match at a jittered offset inside the block
        const auto &p = pv->pats[g.below((uint32_t)pv->pats.size())];
        if ((size_t)p.second <= n) {
            size_t off = (size_t)g.below((uint32_t)(n - (size_t)p.second + 1));
            memcpy(out + off, p.first, (size_t)p.second);
        }
    }
}

}  // namespace

extern "C" {

long long pfac_synth_patterns(int kind, int count, uint64_t seed, int min_len, int max_len,
                              uint8_t *out, size_t cap)
{
    if (count < 0 || min_len < 1 || max_len < min_len || max_len > 1022) return -1;
    Rng g(seed);
    std::unordered_set<std::string> seen;
    std::string all, p;
    long long made = 0;
    int guard = 0;
    while (made < count) {
        p.clear();
        if (kind == PFAC_SYNTH_PAT_PRINTABLE) {   // configs 2 and 4: length uniform, bytes 0x21..0x7E
            int L = g.range(min_len, max_len);
            for (int i = 0; i < L; i++) p += (char)(0x21 + g.below(94));
        } else {                                   // config 3: Snort-like literals
            int L = lognormal_len(g, 12.0, 0.6, min_len, max_len);
            if (g.below(100) < 60) {
                // vocabulary prefix (shared with the text) + a distinctive tail of >= 4 random
                // characters: prefixes are walked often, full matches stay rare (IDS-like)
                p = kPrefixTok[g.below((uint32_t)kNumPrefixTok)];
                if (g.below(4) == 0) p += kPrefixTok[g.below((uint32_t)kNumPrefixTok)];
                int tail = std::max(4, L - (int)p.size());
                if ((int)p.size() + tail > max_len) tail = max_len - (int)p.size();
                for (int i = 0; i < tail; i++) p += kAlnum[g.below((uint32_t)kNumAlnum)];
            } else {
                for (int i = 0; i < L; i++) {
                    uint32_t b = g.below(255);
                    p += (char)(b >= 10 ? b + 1 : b);   // 0x00..0xFF without '\n'
                }
            }
        }
        if (!seen.insert(p).second) {
            if (++guard > 100 * (count + 10)) return -2;   // cannot make that many unique patterns
            continue;
        }
        all += p;
        all += '\n';
        made++;
    }
    if (out && all.size() <= cap) memcpy(out, all.data(), all.size());
    return (long long)all.size();
}

int pfac_synth_text(int kind, uint64_t seed, uint8_t *out, size_t n, const uint8_t *pattern_bytes,
                    size_t pattern_len, int n_threads)
{
    if (!out && n) return -1;
    PatternView pv(pattern_bytes, pattern_bytes ? pattern_len : 0);
    const PatternView *pvp = pattern_bytes ? &pv : nullptr;
    const uint64_t n_blocks = (n + kBlock - 1) / kBlock;
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    if ((uint64_t)n_threads > n_blocks) n_threads = (int)std::max<uint64_t>(1, n_blocks);
    auto work = [&](int tid) {
        for (uint64_t b = (uint64_t)tid; b < n_blocks; b += (uint64_t)n_threads) {
            size_t off = (size_t)b * kBlock;
            fill_block(kind, seed, b, out + off, std::min(kBlock, n - off), pvp);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    return 0;
}

}  // extern "C"
