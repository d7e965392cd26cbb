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

void derive_tables(const Partition &P, uint32_t t2_bytes, uint32_t t3_bytes, uint32_t hot_bytes, Derived &out)
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

    // T1 / T1s over pairs, T2 over 4-byte prefixes: enumerate root paths up to depth 4
    std::vector<uint8_t> t1(65536, 0);
    std::vector<uint32_t> t1s(2048, 0);
    std::vector<uint32_t> t2(t2_bits / 32, 0);
    bool any_short = false;
    std::vector<std::pair<uint32_t, int32_t>> prefix4;   // (4-byte prefix as a little-endian word, state after it)
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
                if ((!t2_bits && t3_bytes < 128) || t2_overflow) continue;
                for (uint32_t e3 = row_begin(s3); e3 < row_end(s3); e3++) {
                    const uint32_t w = pair | ((uint32_t)b2 << 16) | ((uint32_t)edges[e3].byte << 24);
                    if (t2_bits) {
                        const uint32_t h = (w * kHash4Mul) >> out.t2_shift;
                        t2[h >> 5] |= 1u << (h & 31);
                    }
                    prefix4.push_back({w, edges[e3].next});
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

    // ---- Tm / Tm2 / T3: two-point checks (pfac_derive.h).  Level 1: for a 4-byte prefix stored in Tm
    // with m1 = the shortest pattern length below it, the text bytes [m1-4, m1) must be in T3.  Level 2:
    // for a (prefix, level-1 window) group stored in Tm2 with m2 = the shortest length in the group,
    // the text bytes [m2-4, m2) must be in T3 as well.  Keys that are not stored are "unknown" and go
    // straight to the walk, so the filters never reject a start that can match.
    uint32_t t3_bits = 0;
    if (t3_bytes >= 128 && !t2_overflow) {
        t3_bits = 1024;
        while ((uint64_t)t3_bits * 2 <= (uint64_t)t3_bytes * 8) t3_bits *= 2;
    }
    std::vector<uint16_t> tm(kTm1Slots, 0), tm2(kTmSlots, 0);
    std::vector<uint32_t> t3(t3_bits / 32, 0);
    if (t3_bits) {
        int k = 0;
        while ((1u << k) < t3_bits) k++;
        out.t3_shift = 32u - (uint32_t)k;
        // shortest distance from every state to a final state (reverse breadth-first search)
        std::vector<uint32_t> rfirst((size_t)n_states + 2, 0), rsrc(edges.size());
        for (const Edge &e : edges)
            if (e.next < n_states) rfirst[(size_t)e.next + 1]++;
        for (int32_t st = 0; st <= n_states; st++) rfirst[(size_t)st + 1] += rfirst[(size_t)st];
        {
            std::vector<uint32_t> fill(rfirst.begin(), rfirst.end() - 1);
            for (const Edge &e : edges)
                if (e.next < n_states) rsrc[fill[(size_t)e.next]++] = (uint32_t)e.state;
        }
        const uint32_t kInf = 0xFFFFFFFFu;
        std::vector<uint32_t> mind((size_t)n_states, kInf);
        std::vector<int32_t> bfs;
        for (int32_t st = 0; st < std::min(n_final, n_states); st++) { mind[(size_t)st] = 0; bfs.push_back(st); }
        for (size_t i = 0; i < bfs.size(); i++) {
            const int32_t st = bfs[i];
            for (uint32_t e = rfirst[(size_t)st]; e < rfirst[(size_t)st + 1]; e++) {
                const uint32_t src = rsrc[e];
                if (mind[src] == kInf) { mind[src] = mind[(size_t)st] + 1; bfs.push_back((int32_t)src); }
            }
        }
        auto dist = [&](int32_t st) { return st >= 0 && st < n_states ? mind[(size_t)st] : kInf; };

        // 2-choice tagged table: place keys hottest first, then make sure no key that is NOT stored can be
        // mistaken for a stored one (same tag in one of its two slots) -- such stored entries are removed
        struct Key { uint32_t key, m, heat; };
        auto place = [&](std::vector<Key> &keys, std::vector<uint16_t> &tab) -> std::vector<uint8_t> {
            std::stable_sort(keys.begin(), keys.end(), [](const Key &x, const Key &y) { return x.heat > y.heat; });
            std::vector<uint8_t> stored(keys.size(), 0);
            for (size_t i = 0; i < keys.size(); i++) {
                if (keys[i].m < 4 || keys[i].m > 255) continue;
                const uint16_t ent = (uint16_t)((tm_tag(keys[i].key) << 8) | keys[i].m);
                const uint32_t h1 = tm_slot1(keys[i].key), h2 = tm_slot2(keys[i].key);
                // a stored entry with the same tag in either slot would shadow this key: leave it unknown
                if ((tab[h1] && (tab[h1] >> 8) == (ent >> 8)) || (tab[h2] && (tab[h2] >> 8) == (ent >> 8))) continue;
                if (!tab[h1]) { tab[h1] = ent; stored[i] = 1; }
                else if (!tab[h2]) { tab[h2] = ent; stored[i] = 1; }
            }
            bool changed = true;
            while (changed) {
                changed = false;
                for (size_t i = 0; i < keys.size(); i++) {
                    const uint32_t tag = tm_tag(keys[i].key), h1 = tm_slot1(keys[i].key), h2 = tm_slot2(keys[i].key);
                    const uint16_t want = stored[i] ? (uint16_t)((tag << 8) | keys[i].m) : 0;
                    // the kernel takes slot 1 if its tag matches, else slot 2 if its tag matches
                    uint16_t got = 0;
                    if (tab[h1] && (uint32_t)(tab[h1] >> 8) == tag) got = tab[h1];
                    else if (tab[h2] && (uint32_t)(tab[h2] >> 8) == tag) got = tab[h2];
                    if (got != want) {   // this key would read another key's entry: drop that entry
                        if (tab[h1] && (uint32_t)(tab[h1] >> 8) == tag) tab[h1] = 0;
                        else tab[h2] = 0;
                        changed = true;
                    }
                }
                if (changed)   // entries were removed: recompute which keys are still stored
                    for (size_t i = 0; i < keys.size(); i++) {
                        if (!stored[i]) continue;
                        const uint16_t ent = (uint16_t)((tm_tag(keys[i].key) << 8) | keys[i].m);
                        stored[i] = (tab[tm_slot1(keys[i].key)] == ent || tab[tm_slot2(keys[i].key)] == ent) ? 1 : 0;
                    }
            }
            return stored;
        };

        // level 1 keys: one per distinct 4-byte prefix word (a general automaton may reach several states)
        std::sort(prefix4.begin(), prefix4.end());
        std::vector<Key> k1;
        std::vector<std::pair<size_t, size_t>> k1_range;   // prefix4 index range of each key
        for (size_t i = 0; i < prefix4.size();) {
            size_t j = i;
            uint32_t m = kInf, heat_sum = 0;
            while (j < prefix4.size() && prefix4[j].first == prefix4[i].first) {
                const uint32_t d = dist(prefix4[j].second);
                if (d != kInf) m = std::min(m, 4u + d);
                if (prefix4[j].second < n_states) heat_sum += heat[(size_t)prefix4[j].second];
                j++;
            }
            if (m != kInf) {
                k1.push_back({prefix4[i].first, std::min(m, 255u), heat_sum});
                k1_range.push_back({i, j});
            }
            i = j;
        }
        {   // keep ranges aligned with the heat-sorted key order
            std::vector<size_t> idx(k1.size());
            std::iota(idx.begin(), idx.end(), 0);
            std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) { return k1[x].heat > k1[y].heat; });
            std::vector<Key> kk;
            std::vector<std::pair<size_t, size_t>> rr;
            for (size_t x : idx) { kk.push_back(k1[x]); rr.push_back(k1_range[x]); }
            k1.swap(kk);
            k1_range.swap(rr);
        }
        // Level 1 is COMPLETE or absent: every 4-byte prefix gets a slot (cuckoo, 2 buckets x 2 slots), so a
        // miss means "no pattern starts with these 4 bytes" and Tm replaces T2.  Two keys with the same
        // tag that can see each other's entry take the smaller m (any m <= the true minimum is valid).
        std::vector<uint8_t> st1(k1.size(), 0);
        bool complete = k1.size() <= (size_t)kTm1Slots * 17 / 20;
        if (complete) {
            std::vector<int32_t> owner(kTm1Slots, -1);
            auto slots_of = [&](uint32_t key, uint32_t sl[4]) {
                const uint32_t b1 = tm_slot1(key), b2 = tm_slot2(key);
                sl[0] = b1 * 2; sl[1] = b1 * 2 + 1; sl[2] = b2 * 2; sl[3] = b2 * 2 + 1;
            };
            uint64_t rng = 0x9E3779B97F4A7C15ull;
            for (size_t i = 0; i < k1.size() && complete; i++) {
                int32_t cur = (int32_t)i;
                bool placed = false;
                for (int kick = 0; kick < 2000 && !placed; kick++) {
                    uint32_t sl[4];
                    slots_of(k1[(size_t)cur].key, sl);
                    for (int c = 0; c < 4 && !placed; c++)
                        if (owner[sl[c]] < 0) { owner[sl[c]] = cur; placed = true; }
                    if (!placed) {
                        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                        const uint32_t victim = sl[(rng >> 33) & 3];
                        std::swap(cur, owner[victim]);
                    }
                }
                if (!placed) complete = false;
            }
            if (complete) {
                // consistent m among keys the kernel cannot tell apart
                bool changed = true;
                std::vector<uint32_t> mm(k1.size());
                for (size_t i = 0; i < k1.size(); i++) mm[i] = k1[i].m;
                while (changed) {
                    changed = false;
                    for (size_t i = 0; i < k1.size(); i++) {
                        uint32_t sl[4];
                        slots_of(k1[i].key, sl);
                        for (int c = 0; c < 4; c++) {
                            const int32_t o = owner[sl[c]];
                            if (o < 0 || (size_t)o == i || tm_tag(k1[(size_t)o].key) != tm_tag(k1[i].key)) continue;
                            const uint32_t lo = std::min(mm[i], mm[(size_t)o]);
                            if (mm[i] != lo || mm[(size_t)o] != lo) { mm[i] = mm[(size_t)o] = lo; changed = true; }
                        }
                    }
                }
                for (size_t i = 0; i < k1.size(); i++) k1[i].m = mm[i];
                for (uint32_t sl = 0; sl < kTm1Slots; sl++)
                    if (owner[sl] >= 0) tm[sl] = (uint16_t)((tm_tag(k1[(size_t)owner[sl]].key) << 8) | k1[(size_t)owner[sl]].m);
                std::fill(st1.begin(), st1.end(), 1);
                out.tm_complete = 1;
            }
        }
        if (!complete) {   // too many prefixes for the table: no two-point checks, T2 alone filters
            t3_bits = 0;
            t3.clear();
        }

        // walk every path of `steps` bytes below (state, last-4-bytes window); false = too many paths
        uint64_t visited = 0;
        bool overflow = false;
        typedef std::vector<std::pair<int32_t, uint32_t>> Level;
        auto descend = [&](Level &level, uint32_t steps) {
            Level nxt;
            for (uint32_t d = 0; d < steps && !overflow; d++) {
                nxt.clear();
                for (const auto &it : level)
                    for (uint32_t e = row_begin(it.first); e < row_end(it.first); e++) {
                        nxt.push_back({edges[e].next, (it.second >> 8) | ((uint32_t)edges[e].byte << 24)});
                        if (++visited > kDepth4Limit) { overflow = true; break; }
                    }
                level.swap(nxt);
            }
        };
        // level 1 windows -> T3 (seed 1); collect level 2 groups
        struct G2 { uint32_t key2; int32_t state; uint32_t m1; uint32_t w1; };
        std::vector<G2> g2;
        for (size_t i = 0; i < k1.size() && !overflow && complete; i++) {
            if (!st1[i]) continue;
            Level level;
            for (size_t j = k1_range[i].first; j < k1_range[i].second; j++) level.push_back({prefix4[j].second, k1[i].key});
            descend(level, k1[i].m - 4u);
            for (const auto &it : level) {
                const uint32_t h = hash_t3(k1[i].key, it.second) >> out.t3_shift;
                t3[h >> 5] |= 1u << (h & 31);
                g2.push_back({hash_key2(k1[i].key, it.second), it.first, k1[i].m, it.second});
            }
        }
        // level 2 keys: groups by key2; m2 = m1 + shortest distance to a final below any member
        std::sort(g2.begin(), g2.end(), [](const G2 &x, const G2 &y) { return x.key2 < y.key2; });
        std::vector<Key> k2;
        std::vector<std::pair<size_t, size_t>> k2_range;
        for (size_t i = 0; i < g2.size();) {
            size_t j = i;
            uint32_t m = kInf, heat_sum = 0;
            bool same_m1 = true;
            while (j < g2.size() && g2[j].key2 == g2[i].key2) {
                const uint32_t d = dist(g2[j].state);
                if (d != kInf) m = std::min(m, g2[j].m1 + d);
                if (g2[j].m1 != g2[i].m1) same_m1 = false;
                if (g2[j].state < n_states) heat_sum += heat[(size_t)g2[j].state];
                j++;
            }
            // only groups that go deeper than their level-1 window carry information
            if (m != kInf && same_m1 && m > g2[i].m1 && m <= 255u) {
                k2.push_back({g2[i].key2, m, heat_sum});
                k2_range.push_back({i, j});
            } else {
                k2.push_back({g2[i].key2, 0, 0});   // must stay "unknown" (and unshadowed)
                k2_range.push_back({i, j});
            }
            i = j;
        }
        {
            std::vector<size_t> idx(k2.size());
            std::iota(idx.begin(), idx.end(), 0);
            std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) { return k2[x].heat > k2[y].heat; });
            std::vector<Key> kk;
            std::vector<std::pair<size_t, size_t>> rr;
            for (size_t x : idx) { kk.push_back(k2[x]); rr.push_back(k2_range[x]); }
            k2.swap(kk);
            k2_range.swap(rr);
        }
        const std::vector<uint8_t> st2 = place(k2, tm2);
        for (size_t i = 0; i < k2.size() && !overflow; i++) {
            if (!st2[i]) continue;
            for (size_t j = k2_range[i].first; j < k2_range[i].second; j++) {
                Level level(1, {g2[j].state, g2[j].w1});
                descend(level, k2[i].m - g2[j].m1);
                for (const auto &it : level) {
                    const uint32_t h = hash_t3(k2[i].key ^ kT3Seed2, it.second) >> out.t3_shift;
                    t3[h >> 5] |= 1u << (h & 31);
                }
            }
        }
        if (overflow) {   // not a tree: no two-point checks at all
            t3_bits = 0;
            t3.clear();
            out.tm_complete = 0;
        }
        if (t3_bits) {
            for (uint16_t e : tm) out.tm_set += e ? 1u : 0u;
            for (uint16_t e : tm2) out.tm2_set += e ? 1u : 0u;
            out.has_t3 = 1;
        } else {
            out.t3_shift = 32;
        }
    }
    // T2 is only kept when the complete Tm is absent
    if (t3_bits) {
        t2_bits = 0;
        out.t2_shift = 32;
    }
    out.off_t2 = off;
    off = align128(off + t2_bits / 8);
    out.off_tm = off;
    if (t3_bits) off = align128(off + kTm1Slots * 2);
    out.off_tm2 = off;
    if (t3_bits) off = align128(off + kTmSlots * 2);
    out.off_t3 = off;
    off = align128(off + t3_bits / 8);

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
    else t2.clear();
    if (any_short) memcpy(out.image.data() + out.off_t1s, t1s.data(), 8192);
    if (t3_bits) {
        memcpy(out.image.data() + out.off_tm, tm.data(), kTm1Slots * 2);
        memcpy(out.image.data() + out.off_tm2, tm2.data(), kTmSlots * 2);
        memcpy(out.image.data() + out.off_t3, t3.data(), t3_bits / 8);
        for (uint32_t w : t3) out.t3_set += (uint32_t)__builtin_popcount(w);
    }
    if (hot_entries) memcpy(out.image.data() + out.off_hot, hot_tab.data(), (size_t)hot_cap * 8);

    out.val_flagged.resize((size_t)std::max(P.ht_size, 0));
    for (int32_t i = 0; i < P.ht_size; i++) {
        const int32_t v = P.val[(size_t)i];
        out.val_flagged[(size_t)i] = v < 0 ? -1 : (int32_t)flagged(v);
    }
    for (uint8_t b : t1) out.t1_set += b;
    for (uint32_t w : t2) out.t2_set += (uint32_t)__builtin_popcount(w);
}

// Diagnostics only (tools/filter_profile.py): how many start positions of `text` survive each
// stage of the kernel's filter cascade, and how many walk steps the survivors take (split by where
// the step is answered).  Counts only -- no matches are produced here.
void derive_profile(const Partition &P, const Derived &d, const uint8_t *text, size_t n, uint64_t out[12])
{
    const uint8_t *img = d.image.data();
    const uint8_t *t1 = img + d.off_t1;
    const uint32_t *s0f = reinterpret_cast<const uint32_t *>(img + d.off_s0f);
    const uint32_t *t2 = reinterpret_cast<const uint32_t *>(img + d.off_t2);
    const uint32_t *t1s = reinterpret_cast<const uint32_t *>(img + d.off_t1s);
    const uint16_t *tm = reinterpret_cast<const uint16_t *>(img + d.off_tm);
    const uint16_t *tm2 = reinterpret_cast<const uint16_t *>(img + d.off_tm2);
    const uint32_t *t3 = reinterpret_cast<const uint32_t *>(img + d.off_t3);
    const uint32_t *hot = reinterpret_cast<const uint32_t *>(img + d.off_hot);
    auto bit = [](const uint32_t *tab, uint32_t h) { return (tab[h >> 5] >> (h & 31)) & 1u; };
    auto le32 = [&](size_t i) { return (uint32_t)text[i] | ((uint32_t)text[i + 1] << 8) | ((uint32_t)text[i + 2] << 16) | ((uint32_t)text[i + 3] << 24); };
    for (int i = 0; i < 12; i++) out[i] = 0;
    const size_t maxlen = (size_t)std::max(P.max_len, 1);
    for (size_t i = 0; i < n; i++) {
        out[0]++;
        const uint32_t c0 = text[i], c1 = i + 1 < n ? text[i + 1] : 0;
        if (!t1[rot2(c0) | (rot2(c1) << 8)]) continue;
        out[1]++;   // T1 survivors
        bool walk = i + 4 > n;
        if (!walk) {
            const uint32_t w4 = le32(i);
            if (d.has_short && bit(t1s, w4 & 0xffffu)) walk = true;
            else if (d.has_t3 || d.t2_shift >= 32 || bit(t2, (w4 * kHash4Mul) >> d.t2_shift)) {
                out[2]++;   // T2 survivors (everything, when the complete Tm stands in for T2)
                walk = true;
                if (d.has_t3) {
                    const uint32_t m1 = tm1_lookup(tm, w4);
                    if (!m1) { out[3]++; walk = false; }   // no such prefix (Tm is complete)
                    else if (i + m1 > n) walk = false;
                    else {
                        const uint32_t w1 = le32(i + m1 - 4);
                        if (!bit(t3, hash_t3(w4, w1) >> d.t3_shift)) walk = false;
                        else {
                            out[4]++;   // level-1 window passed
                            const uint32_t key2 = hash_key2(w4, w1), m2 = tm_lookup(tm2, key2);
                            if (!m2) out[5]++;   // unknown at level 2
                            else if (i + m2 > n) walk = false;
                            else if (!bit(t3, hash_t3(key2 ^ kT3Seed2, le32(i + m2 - 4)) >> d.t3_shift)) walk = false;
                            else out[6]++;   // level-2 window passed
                        }
                    }
                }
            }
        }
        if (!walk) continue;
        out[7]++;   // starts that walk
        uint32_t sw = s0f[c0];
        size_t q = i + 1;
        const size_t lim = std::min(n, i + maxlen);
        while (sw != kNoState && q < lim) {
            const uint32_t byte = text[q++];
            const uint32_t key = ((sw & d.state_mask) << 8) | byte;
            if (sw & d.hot_bit) {
                out[8]++;   // step answered by the shared-memory hash
                uint32_t slot = d.hot_probe ? (key * d.hot_mul) >> d.hot_shift : 0, nx = kNoState;
                for (uint32_t pr = 0; pr < d.hot_probe; pr++) {
                    out[11]++;   // probes
                    if (hot[slot * 2] == key) { nx = hot[slot * 2 + 1]; break; }
                    if (hot[slot * 2] == kHotEmpty) break;
                    slot = (slot + 1) & d.hot_mask;
                }
                sw = nx;
            } else if ((sw & d.single_bit) && (sw >> 24) != byte) {
                out[9]++;   // step ended by the single-edge look-ahead
                sw = kNoState;
            } else {
                out[10]++;   // step answered by the PHF in L2
                const int32_t nx = P.lookup((int32_t)(sw & d.state_mask), (int32_t)byte);
                if (nx < 0) sw = kNoState;
                else {
                    // the flagged word of nx, as the kernel would read it from {HT,val}
                    const int wb = width_bits(P.width);
                    const int32_t k = (int32_t)key, row = k >> wb;
                    sw = (uint32_t)d.val_flagged[(size_t)(P.r[(size_t)row] + (k & ((1 << wb) - 1)))];
                }
            }
        }
    }
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
    // stage 2 (T1s bypass, T2, Tm/T3, Tm2/T3) must pass every pattern: run the kernel's filter logic
    // over each pattern's own bytes (strings of the breadth-first tree from the root row)
    {
        const uint16_t *tm = reinterpret_cast<const uint16_t *>(img + d.off_tm);
        const uint16_t *tm2 = reinterpret_cast<const uint16_t *>(img + d.off_tm2);
        const uint32_t *t3 = reinterpret_cast<const uint32_t *>(img + d.off_t3);
        const int32_t n_states = std::max(P.state_num, 0);
        std::vector<int32_t> par((size_t)n_states, -2), order;
        std::vector<uint8_t> pbyte((size_t)n_states, 0);
        for (int b = 0; b < kCharSet; b++) {
            const int32_t s1 = P.s0.empty() ? -1 : P.s0[(size_t)b];
            if (s1 >= 0 && s1 < n_states && par[(size_t)s1] == -2) { par[(size_t)s1] = -1; pbyte[(size_t)s1] = (uint8_t)b; order.push_back(s1); }
        }
        for (size_t i = 0; i < order.size(); i++)
            for (int b = 0; b < kCharSet; b++) {
                const int32_t y = P.lookup(order[i], b);
                if (y >= 0 && y < n_states && par[(size_t)y] == -2) { par[(size_t)y] = order[i]; pbyte[(size_t)y] = (uint8_t)b; order.push_back(y); }
            }
        auto le32 = [](const uint8_t *q) { return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24); };
        auto bit = [](const uint32_t *tab, uint32_t h) { return (tab[h >> 5] >> (h & 31)) & 1u; };
        std::vector<uint8_t> str;
        for (int32_t f = 0; f < std::min(P.n_final, n_states); f++) {
            if (par[(size_t)f] == -2) continue;   // unreachable final (duplicate pattern)
            str.clear();
            for (int32_t x = f; x >= 0; x = par[(size_t)x]) str.push_back(pbyte[(size_t)x]);
            std::reverse(str.begin(), str.end());
            const uint32_t len = (uint32_t)str.size();
            if (len < 4) continue;   // covered by the T1s check above
            const uint32_t w4 = le32(str.data());
            const uint32_t pair = w4 & 0xffffu;
            if (d.has_short && bit(t1s, pair)) continue;
            if (!d.has_t3) {
                if (d.t2_shift < 32 && !bit(t2, (w4 * kHash4Mul) >> d.t2_shift)) return 6;
                continue;
            }
            const uint32_t m1 = tm1_lookup(tm, w4);
            if (!m1) return 20;   // Tm is complete: every prefix must be found
            if (m1 > len) return 16;
            const uint32_t w1 = le32(str.data() + m1 - 4);
            if (!bit(t3, hash_t3(w4, w1) >> d.t3_shift)) return 17;
            const uint32_t key2 = hash_key2(w4, w1);
            const uint32_t m2 = tm_lookup(tm2, key2);
            if (!m2) continue;
            if (m2 > len) return 18;
            const uint32_t w2 = le32(str.data() + m2 - 4);
            if (!bit(t3, hash_t3(key2 ^ kT3Seed2, w2) >> d.t3_shift)) return 19;
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
