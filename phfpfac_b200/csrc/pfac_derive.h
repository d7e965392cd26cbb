// Device-side derived layouts of one partition's canonical tables (host code, no CUDA).
//
// The canonical PHF arrays r/HT/val (bit-compatible with CreateTable/FFDM, reference phf.c:151)
// stay the authoritative transition function and are uploaded unchanged (HT and val interleaved).
// What is derived here are the shared-memory resident ACCELERATORS the scan kernel consults
// first; every one of them is a superset filter or an exact copy of PHF rows, so lookups stay
// equivalent to master_kernel.cu:52-64:
//   T1    65,536 x u8 : byte 1 iff a walk that starts with bytes (c0,c1) matches a 1-byte pattern
//                       or has a second edge  (root fan-out, s0Table of main.cc:200, folded in)
//   T1s   65,536 bits : pair (c0,c1) can complete a pattern of length <= 3 (bypasses T2)
//   T2    2^k2 bits   : multiplicative hash of every 4-byte pattern prefix
//   Tm    8192 x u16  : COMPLETE cuckoo table (2 buckets x 2 tagged slots) of every 4-byte prefix
//                       -> m1 = shortest pattern length below it; a miss rejects the start, so Tm
//                       stands in for T2 (sets with too many prefixes keep T2 and get no Tm/T3)
//   Tm2   4096 x u16  : same, (prefix, level-1 window) group -> m2 = shortest length in the group
//   T3    2^k3 bits   : hash of (key, pattern bytes [m-4, m)) for every pattern below a stored key:
//                       a start whose text at offset m-4 is not in T3 cannot complete any pattern
//                       under that key (Wu-Manber style two-point checks; they end the walks along
//                       prefixes the text shares with many patterns).  Keys that are not stored
//                       are "unknown": the start simply walks.
//   s0f   256 x u32   : root row as state words
//   hot   open-addressing hash of COMPLETE PHF rows of the hottest states
//                       (key = state<<8|byte -> next state word); a miss in a hot row means
//                       "no transition", exactly like HT[idx] != row in the PHF
// State words (s0f, hot values, the val half of the global {HT,val} array) carry a look-ahead on
// the TARGET state's own row, so most walks end without touching L2:
//   bits 0..21  state number
//   bit  22     hot: the state's row is answered by the shared-memory hash (a leaf, i.e. a row
//               with no transition at all, is hot by definition and needs no table entry)
//   bit  23     single: the row has exactly one transition, on the byte in bits 24..31 -- any other
//               byte ends the walk without a lookup
// Automata with 2^22 states or more get plain words (no flags, no hot table).
#pragma once
#include <cstdint>
#include <vector>

#include "pfac_internal.h"

namespace pfac {

constexpr uint32_t kStateBits = 22;
constexpr uint32_t kHotFlag = 1u << 22;
constexpr uint32_t kSingleFlag = 1u << 23;
constexpr uint32_t kStateMask = (1u << kStateBits) - 1;
constexpr uint32_t kNoState = 0xFFFFFFFFu;
constexpr uint32_t kHotEmpty = 0xFFFFFFFFu;
constexpr uint32_t kHash4Mul = 0x9E3779B1u;
constexpr uint32_t kTmSlotBits = 12;
constexpr uint32_t kTmSlots = 1u << kTmSlotBits;   // u16 entries: tag << 8 | m, 0 = empty; two choices per key
constexpr uint32_t kTm1Slots = 2 * kTmSlots;       // level 1: 4096 buckets x 2 slots, cuckoo, complete
constexpr uint32_t kT3Seed2 = 0x5bd1e995u;

#if defined(__CUDACC__)
#define PFAC_HD __host__ __device__
#else
#define PFAC_HD
#endif
// T3 index (before the final shift) of a key and the 4 window bytes (little-endian word)
PFAC_HD inline uint32_t hash_t3(uint32_t key, uint32_t window)
{
    uint32_t x = key * 0x9E3779B1u + window * 0x85EBCA6Bu;
    x ^= x >> 15;
    return x * 0x2C1B3C6Du;
}
// level-2 key of a (4-byte prefix, level-1 window) group
PFAC_HD inline uint32_t hash_key2(uint32_t prefix, uint32_t window)
{
    uint32_t x = prefix * 0xC2B2AE35u ^ window * 0x27D4EB2Fu;
    x ^= x >> 16;
    return x * 0x165667B1u + 0x9E3779B9u;
}
PFAC_HD inline uint32_t tm_slot1(uint32_t key) { return (key * 0xC2B2AE35u) >> (32 - kTmSlotBits); }
PFAC_HD inline uint32_t tm_slot2(uint32_t key) { return (key * 0x27D4EB2Fu + 0x7F4A7C15u) >> (32 - kTmSlotBits); }
PFAC_HD inline uint32_t tm_tag(uint32_t key) { return (key * 0xFD7046C5u) >> 24; }
// level 1 (complete table, 2 buckets x 2 slots): m of `key`, 0 = no pattern has this 4-byte prefix
PFAC_HD inline uint32_t tm1_lookup(const uint16_t *tab, uint32_t key)
{
    const uint32_t tag = tm_tag(key);
    const uint32_t *t32 = reinterpret_cast<const uint32_t *>(tab);
    const uint32_t a = t32[tm_slot1(key)], b = t32[tm_slot2(key)];
    if ((a & 0xffffu) && ((a >> 8) & 0xffu) == tag) return a & 255u;
    if ((a >> 16) && (a >> 24) == tag) return (a >> 16) & 255u;
    if ((b & 0xffffu) && ((b >> 8) & 0xffu) == tag) return b & 255u;
    if ((b >> 16) && (b >> 24) == tag) return (b >> 16) & 255u;
    return 0;
}
// the entry the kernel uses for `key` (0 = unknown): slot 1 if its tag matches, else slot 2
PFAC_HD inline uint32_t tm_lookup(const uint16_t *tab, uint32_t key)
{
    const uint32_t tag = tm_tag(key);
    const uint32_t e1 = tab[tm_slot1(key)], e2 = tab[tm_slot2(key)];
    if (e1 && (e1 >> 8) == tag) return e1 & 255u;
    if (e2 && (e2 >> 8) == tag) return e2 & 255u;
    return 0;
}

struct Derived {
    // shared-memory image, copied verbatim by the kernel (sections 128-byte aligned)
    std::vector<uint8_t> image;
    uint32_t off_t1 = 0, off_s0f = 0, off_t2 = 0, off_t1s = 0, off_tm = 0, off_tm2 = 0, off_t3 = 0, off_hot = 0;
    uint32_t has_t3 = 0, t3_shift = 32, t3_set = 0, tm_set = 0, tm2_set = 0, tm_complete = 0;
    uint32_t t2_shift = 32;      // index = (w * kHash4Mul) >> t2_shift
    uint32_t has_short = 0;      // patterns of length <= 3 exist (T1s present)
    uint32_t state_mask = 0x7FFFFFFFu, hot_bit = 0, single_bit = 0;   // plain words unless flags fit
    uint32_t hot_mask = 0;       // entries - 1 (0: no hot table)
    uint32_t hot_shift = 32;     // slot = (key * hot_mul) >> hot_shift
    uint32_t hot_mul = 0;
    uint32_t hot_probe = 0;      // longest probe sequence needed
    uint32_t n_hot_rows = 0, n_hot_entries = 0;
    // global-memory copy of val with the hot flag (HT interleaved by the uploader)
    std::vector<int32_t> val_flagged;
    // statistics (pfac_ctx_derived_info)
    uint32_t t1_set = 0, t2_set = 0, n_depth4 = 0;
};

// rotl2 of every byte: the kernel indexes T1 with text bytes rotated left by 2 so that the low,
// high-entropy bits of ASCII text select the shared-memory bank
inline uint32_t rot2(uint32_t c) { return ((c << 2) | (c >> 6)) & 0xFFu; }

// t2_bytes / hot_bytes: shared-memory budget of the two variable sections (powers of two; 0 = none)
void derive_tables(const Partition &P, uint32_t t2_bytes, uint32_t t3_bytes, uint32_t hot_bytes, Derived &out);

int derive_selfcheck(const Partition &P, const Derived &d);
void derive_profile(const Partition &P, const Derived &d, const uint8_t *text, size_t n, uint64_t out[12]);

}  // namespace pfac
