// Host-side derivation of the scan kernel's shared-memory accelerators from one partition's
// canonical PHF arrays (see pfac_derive.h).  Everything here is computed FROM r/HT/val/s0Table,
// i.e. from what CreateTable + FFDM (reference create_table_reorder.c:277, phf.c:151) emit, so
// tables handed in through pfac_tables_from_arrays get the same treatment.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "pfac_derive.h"

namespace pfac {

namespace {

struct Edge {
    int32_t state, byte, next;
};

inline uint32_t align128(uint32_t x) { return (x + 127u) & ~127u; }

}  // namespace

void derive_tables(const Partition &P, uint32_t t2_bytes, uint32_t hot_bytes, Derived &out)
{
    out = Derived();
    const int32_t n_states = std::max(P.state_num, 0);
    const int32_t n_final = P.n_final;
    const int wb = width_bits(P.width);

    // ---- transitions, recovered from the PHF slots: slot idx holds key row*width + (idx - r[row])
    std::vector<Edge> edges;
    edges.reserve((size_t)std::max(P.n_keys, 0));
    for (int32_t idx = 0; idx < P.ht_size; idx++) {
        const int32_t row = P.HT[(size_t)idx];
        if (row < 0 || row >= (int32_t)P.r.size()) continue;
        const int64_t col = (int64_t)idx - P.r[(size_t)row];
        if (col < 0 || col >= P.width) continue;
        const int64_t key = ((int64_t)row << wb) + col;
        const int64_t st = key >> 8;
        const int32_t nx = P.val[(size_t)idx];
        if (st >= n_states || nx < 0) continue;   // never produced by a lookup the kernel can make
        edges.push_back({(int32_t)st, (int32_t)(key & 255), nx});
    }
    std::sort(edges.begin(), edges.end(), [](const Edge &a, const Edge &b) {
        return a.state != b.state ? a.state < b.state : a.byte < b.byte;
    });
    std::vector<uint32_t> first((size_t)n_states + 2, 0);
    for (const Edge &e : edges) first[(size_t)e.state + 1]++;
    for (int32_t s = 0; s <= n_states; s++) first[(size_t)s + 1] += first[(size_t)s];
    auto row_begin = [&](int32_t s) { return s >= 0 && s < n_states ? first[(size_t)s] : 0u; };
    auto row_end = [&](int32_t s) { return s >= 0 && s < n_states ? first[(size_t)s + 1] : 0u; };
    auto is_final = [&](int32_t s) { return s >= 0 && s < n_final; };

    // ---- breadth-first tree from the root row (first visit = parent); heat = patterns below
    std::vector<int32_t> order, parent((size_t)n_states, -2), depth((size_t)n_states, 0);
    order.reserve((size_t)n_states);
    for (int b = 0; b < kCharSet; b++) {
        const int32_t s = P.s0.empty() ? -1 : P.s0[(size_t)b];
        if (s >= 0 && s < n_states && parent[(size_t)s] == -2) {
            parent[(size_t)s] = -1;
            depth[(size_t)s] = 1;
            order.push_back(s);
        }
    }
    for (size_t i = 0; i < order.size(); i++) {
        const int32_t s = order[i];
        for (uint32_t e = row_begin(s); e < row_end(s); e++) {
            const int32_t t = edges[e].next;
            if (t < n_states && parent[(size_t)t] == -2) {
                parent[(size_t)t] = s;
                depth[(size_t)t] = depth[(size_t)s] + 1;
                order.push_back(t);
            }
        }
    }
    std::vector<uint32_t> heat((size_t)n_states, 0);
    for (size_t i = order.size(); i-- > 0;) {
        const int32_t s = order[i];
        heat[(size_t)s] += is_final(s) ? 1u : 0u;
        if (parent[(size_t)s] >= 0) heat[(size_t)parent[(size_t)s]] += heat[(size_t)s];
    }

    // ---- hot rows: complete PHF rows of the hottest states, as many as the budget holds
    uint32_t hot_entries = 0;
    std::vector<uint8_t> is_hot((size_t)n_states, 0);
    std::vector<int32_t> hot_rows;
    const bool key_fits = n_states < (1 << kStateBits);   // room for the flags above the state number
    if (key_fits) {
        out.state_mask = kStateMask;
        out.hot_bit = kHotFlag;
        out.single_bit = kSingleFlag;
    }
    uint32_t hot_cap = 0;
    if (hot_bytes >= 1024 && key_fits && !order.empty()) {
        hot_cap = 1;
        while ((uint64_t)hot_cap * 2 * 8 <= hot_bytes) hot_cap *= 2;
        const uint32_t budget = (uint32_t)((uint64_t)hot_cap * 11 / 20);   // load factor <= 0.55
        std::vector<int32_t> cand;
        for (int32_t s : order)
            if (row_end(s) > row_begin(s)) cand.push_back(s);
        std::stable_sort(cand.begin(), cand.end(), [&](int32_t a, int32_t b) {
            if (heat[(size_t)a] != heat[(size_t)b]) return heat[(size_t)a] > heat[(size_t)b];
            return depth[(size_t)a] < depth[(size_t)b];
        });
        for (int32_t s : cand) {
            const uint32_t n = row_end(s) - row_begin(s);
            if (heat[(size_t)s] < 2) break;          // a single pattern below: never "hot"
            if (hot_entries + n > budget) continue;  // a narrower, cooler row may still fit
            is_hot[(size_t)s] = 1;
            hot_rows.push_back(s);
            hot_entries += n;
        }
    }
    auto flagged = [&](int32_t s) -> uint32_t {
        if (s < 0) return kNoState;
        if (!key_fits || s >= n_states) return (uint32_t)s;
        const uint32_t n = row_end(s) - row_begin(s);
        if (n == 0 || is_hot[(size_t)s]) return (uint32_t)s | kHotFlag;   // a leaf is a complete (empty) hot row
        if (n == 1) return (uint32_t)s | kSingleFlag | ((uint32_t)edges[row_begin(s)].byte << 24);
        return (uint32_t)s;
    };

    // ---- image layout
    uint32_t off = 0;
    out.off_t1 = off;
    off = align128(off + 65536);
    out.off_s0f = off;
    off = align128(off + 1024);
    uint32_t t2_bits = 0;
    if (t2_bytes >= 128) {
        t2_bits = 1024;
        while ((uint64_t)t2_bits * 2 <= (uint64_t)t2_bytes * 8) t2_bits *= 2;
    }
    out.off_t2 = off;
    off = align128(off + t2_bits / 8);

    // T1 / T1s over pairs, T2 over 4-byte prefixes: enumerate root paths up to depth 4
    std::vector<uint8_t> t1(65536, 0);
    std::vector<uint32_t> t1s(2048, 0);
    std::vector<uint32_t> t2(t2_bits / 32, 0);
    bool any_short = false;
    uint64_t n_depth4 = 0;
    bool t2_overflow = false;
    const uint64_t kDepth4Limit = 1ull << 23;
    if (t2_bits) {
        int k = 0;
        while ((1u << k) < t2_bits) k++;
        out.t2_shift = 32u - (uint32_t)k;
    }
    for (int b0 = 0; b0 < kCharSet; b0++) {
        const int32_t s1 = P.s0.empty() ? -1 : P.s0[(size_t)b0];
        if (s1 < 0) continue;
        if (is_final(s1)) {   // a 1-byte pattern: every pair starting with b0 reports a match
            any_short = true;
            for (int b1 = 0; b1 < kCharSet; b1++) {
                t1[(rot2((uint32_t)b0)) | (rot2((uint32_t)b1) << 8)] = 1;
                const uint32_t pair = (uint32_t)b0 | ((uint32_t)b1 << 8);
                t1s[pair >> 5] |= 1u << (pair & 31);
            }
        }
        for (uint32_t e1 = row_begin(s1); e1 < row_end(s1); e1++) {
            const int32_t b1 = edges[e1].byte, s2 = edges[e1].next;
            const uint32_t pair = (uint32_t)b0 | ((uint32_t)b1 << 8);
            t1[rot2((uint32_t)b0) | (rot2((uint32_t)b1) << 8)] = 1;
            bool shortp = is_final(s2);
            for (uint32_t e2 = row_begin(s2); e2 < row_end(s2); e2++) {
                const int32_t b2 = edges[e2].byte, s3 = edges[e2].next;
                if (is_final(s3)) shortp = true;
                if (!t2_bits || t2_overflow) continue;
                for (uint32_t e3 = row_begin(s3); e3 < row_end(s3); e3++) {
                    const uint32_t w = pair | ((uint32_t)b2 << 16) | ((uint32_t)edges[e3].byte << 24);
                    const uint32_t h = (w * kHash4Mul) >> out.t2_shift;
                    t2[h >> 5] |= 1u << (h & 31);
                    if (++n_depth4 > kDepth4Limit) { t2_overflow = true; break; }
                }
            }
            if (shortp) {
                any_short = true;
                t1s[pair >> 5] |= 1u << (pair & 31);
            }
        }
    }
    if (t2_overflow) std::fill(t2.begin(), t2.end(), 0xFFFFFFFFu);   // not a tree: T2 filters nothing
    out.has_short = any_short ? 1u : 0u;
    out.n_depth4 = (uint32_t)std::min<uint64_t>(n_depth4, 0xFFFFFFFFu);
    out.off_t1s = off;
    if (any_short) off = align128(off + 8192);

    // hot hash: open addressing, linear probing; pick the multiplier with the shortest probes
    out.off_hot = off;
    std::vector<uint32_t> hot_tab;
    if (hot_entries) {
        int k = 0;
        while ((1u << k) < hot_cap) k++;
        out.hot_mask = hot_cap - 1;
        out.hot_shift = 32u - (uint32_t)k;
        static const uint32_t muls[] = {0x9E3779B1u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu, 0x165667B1u,
                                        0xD3A2646Du, 0xFD7046C5u, 0xB55A4F09u};
        uint32_t best_probe = 0xFFFFFFFFu;
        std::vector<uint32_t> tab;
        for (uint32_t mul : muls) {
            tab.assign((size_t)hot_cap * 2, kHotEmpty);
            uint32_t worst = 0;
            for (int32_t s : hot_rows) {
                for (uint32_t e = row_begin(s); e < row_end(s); e++) {
                    const uint32_t key = ((uint32_t)s << 8) | (uint32_t)edges[e].byte;
                    uint32_t slot = (key * mul) >> out.hot_shift, probes = 1;
                    while (tab[(size_t)slot * 2] != kHotEmpty) {
                        slot = (slot + 1) & out.hot_mask;
                        probes++;
                    }
                    tab[(size_t)slot * 2] = key;
                    tab[(size_t)slot * 2 + 1] = flagged(edges[e].next);
                    worst = std::max(worst, probes);
                }
            }
            // a miss probes until the first empty slot: measure the longest occupied run too
            uint32_t run = 0, longest = 0;
            for (uint32_t i = 0; i < 2 * hot_cap; i++) {
                if (tab[(size_t)(i & out.hot_mask) * 2] != kHotEmpty) { run++; longest = std::max(longest, run); }
                else run = 0;
                if (run >= hot_cap) break;
            }
            worst = std::max(worst, longest + 1);
            if (worst < best_probe) {
                best_probe = worst;
                out.hot_mul = mul;
                hot_tab = tab;
            }
        }
        out.hot_probe = best_probe;
        out.n_hot_rows = (uint32_t)hot_rows.size();
        out.n_hot_entries = hot_entries;
        off = align128(off + hot_cap * 8);
    } else {
        std::fill(is_hot.begin(), is_hot.end(), 0);
    }

    out.image.assign(off, 0);
    memcpy(out.image.data() + out.off_t1, t1.data(), 65536);
    uint32_t *s0f = reinterpret_cast<uint32_t *>(out.image.data() + out.off_s0f);
    for (int b = 0; b < kCharSet; b++) s0f[b] = flagged(P.s0.empty() ? -1 : P.s0[(size_t)b]);
    if (t2_bits) memcpy(out.image.data() + out.off_t2, t2.data(), t2_bits / 8);
    if (any_short) memcpy(out.image.data() + out.off_t1s, t1s.data(), 8192);
    if (hot_entries) memcpy(out.image.data() + out.off_hot, hot_tab.data(), (size_t)hot_cap * 8);

    out.val_flagged.resize((size_t)std::max(P.ht_size, 0));
    for (int32_t i = 0; i < P.ht_size; i++) {
        const int32_t v = P.val[(size_t)i];
        out.val_flagged[(size_t)i] = v < 0 ? -1 : (int32_t)flagged(v);
    }
    for (uint8_t b : t1) out.t1_set += b;
    for (uint32_t w : t2) out.t2_set += (uint32_t)__builtin_popcount(w);
}

// Host model of what the kernel does with the derived tables, checked against the canonical PHF:
// returns 0 if (a) every root path passes T1, T1s covers every pair that can complete a pattern
// of <= 3 bytes and T2 holds every 4-byte prefix, and (b) every hot row answers exactly like
// Partition::lookup for all 256 bytes.  Non-zero = index of the first violated invariant.
int derive_selfcheck(const Partition &P, const Derived &d)
{
    const uint8_t *img = d.image.data();
    if (d.image.empty()) return 100;
    const uint8_t *t1 = img + d.off_t1;
    const uint32_t *s0f = reinterpret_cast<const uint32_t *>(img + d.off_s0f);
    const uint32_t *t2 = reinterpret_cast<const uint32_t *>(img + d.off_t2);
    const uint32_t *t1s = reinterpret_cast<const uint32_t *>(img + d.off_t1s);
    const uint32_t *hot = reinterpret_cast<const uint32_t *>(img + d.off_hot);
    auto hot_lookup = [&](uint32_t sw, uint32_t byte) -> uint32_t {
        const uint32_t key = ((sw & d.state_mask) << 8) | byte;
        uint32_t slot = d.hot_probe ? (key * d.hot_mul) >> d.hot_shift : 0;
        for (uint32_t pr = 0; pr < d.hot_probe; pr++) {
            if (hot[slot * 2] == key) return hot[slot * 2 + 1];
            if (hot[slot * 2] == kHotEmpty) break;
            slot = (slot + 1) & d.hot_mask;
        }
        return kNoState;
    };
    auto is_final = [&](int32_t s) { return s >= 0 && s < P.n_final; };
    std::vector<uint32_t> all_words;   // every state word the kernel can hold
    for (int b0 = 0; b0 < kCharSet; b0++) {
        const int32_t s1 = P.s0.empty() ? -1 : P.s0[(size_t)b0];
        if ((s1 < 0) != (s0f[b0] == kNoState)) return 1;
        if (s1 < 0) continue;
        if ((s0f[b0] & d.state_mask) != (uint32_t)s1) return 2;
        all_words.push_back(s0f[b0]);
        for (int b1 = 0; b1 < kCharSet; b1++) {
            const int32_t s2 = P.lookup(s1, b1);
            const uint32_t pair = (uint32_t)b0 | ((uint32_t)b1 << 8);
            const bool pass = t1[rot2((uint32_t)b0) | (rot2((uint32_t)b1) << 8)] != 0;
            if ((is_final(s1) || s2 >= 0) && !pass) return 3;
            if (!(is_final(s1) || s2 >= 0) && pass) return 4;   // T1 is exact, not just a superset
            bool shortp = is_final(s1) || is_final(s2);
            if (s2 < 0) { if (shortp && !d.has_short) return 5; continue; }
            for (int b2 = 0; b2 < kCharSet; b2++) {
                const int32_t s3 = P.lookup(s2, b2);
                if (s3 < 0) continue;
                if (is_final(s3)) shortp = true;
                for (int b3 = 0; b3 < kCharSet; b3++) {
                    if (P.lookup(s3, b3) < 0) continue;
                    const uint32_t w = pair | ((uint32_t)b2 << 16) | ((uint32_t)b3 << 24);
                    if (d.t2_shift < 32) {
                        const uint32_t h = (w * kHash4Mul) >> d.t2_shift;
                        if (!((t2[h >> 5] >> (h & 31)) & 1u)) return 6;
                    }
                }
            }
            if (shortp && !(d.has_short && ((t1s[pair >> 5] >> (pair & 31)) & 1u))) return 7;
        }
    }
    // every state word must describe its state's row truthfully (hot rows complete, single-edge
    // byte right, leaves flagged hot); words come from s0f, the hot values and the flagged val[]
    std::vector<uint8_t> seen((size_t)std::max(P.state_num, 1), 0);
    for (int32_t i = 0; i < P.ht_size; i++) {
        const int32_t v = d.val_flagged[(size_t)i];
        if ((v == -1) != (P.val[(size_t)i] < 0)) return 9;
        if (P.val[(size_t)i] < 0) continue;
        if (((uint32_t)v & d.state_mask) != (uint32_t)P.val[(size_t)i]) return 8;
        all_words.push_back((uint32_t)v);
    }
    uint32_t rows = 0;
    for (size_t i = 0; i < all_words.size(); i++) {
        const uint32_t sw = all_words[i];
        const int32_t s = (int32_t)(sw & d.state_mask);
        if (s >= P.state_num) { if (sw != (uint32_t)s) return 10; continue; }
        int n_edges = 0, only = -1;
        if (sw & (d.hot_bit | d.single_bit))
            for (int b = 0; b < kCharSet; b++)
                if (P.lookup(s, b) >= 0) { n_edges++; only = b; }
        if (sw & d.single_bit) {
            if ((sw & d.hot_bit) || n_edges != 1 || (int)(sw >> 24) != only) return 15;
        }
        if (!(sw & d.hot_bit) || seen[(size_t)s]) continue;
        seen[(size_t)s] = 1;
        if (n_edges) rows++;
        for (int b = 0; b < kCharSet; b++) {
            const int32_t want = P.lookup(s, b);
            const uint32_t got = hot_lookup(sw, (uint32_t)b);
            if ((want < 0) != (got == kNoState)) return 12;
            if (want >= 0 && (got & d.state_mask) != (uint32_t)want) return 13;
            if (want >= 0) all_words.push_back(got);
        }
    }
    if (rows > d.n_hot_rows) return 14;
    return 0;
}

}  // namespace pfac
