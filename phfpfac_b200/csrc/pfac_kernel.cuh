// PFAC scan kernel for sm_100a (B200).  Replaces TraceTable_kernel + SUBSEG_MATCH
// (reference master_kernel.cu:37-180).
//
// Design:
//   * persistent CTAs pull tile tickets from an atomic counter; every tile (start positions
//     plus a halo of max_pat_len-1 bytes) is staged into shared memory by one
//     cp.async.bulk (TMA bulk copy, SASS UBLKCP) per tile, double buffered on mbarriers;
//   * phase 1 (filter): every lane tests 16 start positions against a shared-memory bitmap of
//     all 2-byte pattern prefixes (root fan-out folded in); survivors are compacted, in
//     position order, into a per-warp queue (prefix-popc);
//   * phase 2 (walk): full warps walk the queued starts through the PHF (r[] then the
//     interleaved {HT,val} slot, read-only L2/L1-resident loads) and count matches;
//   * phase 3 (emit): tile totals go through a decoupled look-back over tile tickets, so the
//     (pos,id) records land in global memory ordered by position without a second pass over
//     the input; only starts that matched are walked again to write their records.
// Output order within a start position is walk depth = pattern length ascending, the order
// main.cc:341-349 prints.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfac {

struct ScanParams {
    const uint8_t *in_al;     // 16-byte aligned base: caller's pointer rounded down
    uint32_t mis;             // offset of the first start position in the aligned stream (0..15)
    uint32_t a_start_end;     // mis + n_starts   (aligned-stream coordinates, exclusive)
    uint32_t a_valid_end;     // mis + n_valid
    uint32_t n_tiles;
    uint32_t halo;            // staged halo bytes, multiple of 16, >= max_pat_len-1
    uint32_t max_pat_len;
    int32_t use_ref_bound;    // reproduce the 4096+512 walk bound (master_kernel.cu:141-144)
    uint64_t base_pos;        // global position of the first start position
    uint32_t pos_bias;        // added to every record position (sub-chunk offset inside a host call)
    const int32_t *r;         // canonical r[]                      (phf.c:197)
    const int2 *htval;        // {HT[i], val[i]} interleaved        (phf.c:211,216)
    const int32_t *idmap;     // final state -> pattern id          (create_table_reorder.c:318)
    const int32_t *s0;        // root row                           (main.cc:200)
    const uint32_t *bitmap2;  // 65536 bits: bit (b0 | b1<<8) set iff a walk from b0,b1 can go on or match
    int32_t ht_size, width_bit, n_final;
    uint2 *out;               // pfac_match records
    unsigned long long cap;
    unsigned long long *count_out;
    unsigned long long *tile_state;   // decoupled look-back: [63:62] status, [61:0] value
    unsigned int *ticket;
    unsigned int *error_flag;
    uint32_t debug;           // PFAC_DEBUG bits (timing experiments only): 1 no walk, 2 no look-back, 4 no filter
};

constexpr int kThreads = 512;          // 16 warps per CTA
constexpr int kWarpRange = 1024;       // start positions per warp per tile (2 steps of 32 lanes x 16 B)
constexpr int kWarps = kThreads / 32;
constexpr int kTile = kWarps * kWarpRange;   // 16 KiB of start positions per tile
constexpr int kSmemFixed = 9472;       // bitmap 8192 + s0 1024 + control 256
constexpr unsigned kSpinLimit = 1u << 26;

__host__ __device__ inline uint32_t scan_buf_stride(uint32_t halo) { return (kTile + halo + 32 + 127) & ~127u; }
__host__ inline size_t scan_smem_bytes(uint32_t halo)
{
    return (size_t)kSmemFixed + (size_t)kWarps * kWarpRange * 2 + 2 * (size_t)scan_buf_stride(halo);
}

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr unsigned long long kStA = 1ull << 62;   // tile aggregate published
constexpr unsigned long long kStP = 2ull << 62;   // inclusive prefix published
constexpr unsigned long long kStMask = (1ull << 62) - 1;

// Decoupled look-back over tile tickets (one warp).  Returns the number of matches in all
// tiles before `tile`; publishes this tile's aggregate first and its inclusive prefix last.
// Tickets are handed out in order, so every predecessor is held by a running CTA.
__device__ __forceinline__ unsigned long long tile_lookback(unsigned long long *st, uint32_t tile,
                                                            unsigned long long T, int lane,
                                                            unsigned int *error_flag)
{
    if (tile == 0) {
        if (lane == 0) st_relaxed_gpu(&st[0], kStP | T);
        return 0;
    }
    if (lane == 0) st_relaxed_gpu(&st[tile], kStA | T);
    unsigned long long excl = 0;
    long long j = (long long)tile - 1;
    unsigned spins = 0;
    while (true) {
        long long idx = j - lane;
        unsigned long long s = idx >= 0 ? ld_relaxed_gpu(&st[idx]) : kStP;
        unsigned status = (unsigned)(s >> 62);
        unsigned inval = __ballot_sync(0xffffffffu, status == 0);
        unsigned pm = __ballot_sync(0xffffffffu, status == 2);
        int firstP = pm ? (__ffs(pm) - 1) : 32;
        unsigned need = firstP >= 31 ? 0xffffffffu : ((2u << firstP) - 1u);
        if (inval & need) {
            if (++spins > kSpinLimit) {   // watchdog: never expected
                if (lane == 0) atomicExch(error_flag, 1u);
                break;
            }
            __nanosleep(64);
            continue;
        }
        unsigned long long v = (lane <= firstP) ? (s & kStMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (firstP < 32) break;
        j -= 32;
    }
    if (lane == 0) st_relaxed_gpu(&st[tile], kStP | ((excl + T) & kStMask));
    return excl;
}

// 16 start positions per lane: bit i of the result is set iff the 2-byte window at byte i
// is in the prefix bitmap.
__device__ __forceinline__ uint32_t filter16(const uint4 v, const uint32_t nx, const uint32_t *__restrict__ bm)
{
    const uint32_t w[5] = {v.x, v.y, v.z, v.w, nx};
    uint32_t mask = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t win;
            if (i == 0) win = w[k] & 0xffffu;
            else if (i == 1) win = (w[k] >> 8) & 0xffffu;
            else if (i == 2) win = w[k] >> 16;
            else win = __funnelshift_r(w[k], w[k + 1], 24) & 0xffffu;
            uint32_t bit = (bm[win >> 5] >> (win & 31u)) & 1u;
            mask |= bit << (4 * k + i);
        }
    }
    return mask;
}

// One start position: the walk of SUBSEG_MATCH (master_kernel.cu:39-73).  tpos = tile-relative
// byte index of the start, lim_t = tile-relative exclusive bound of readable bytes.
// WRITE == false: returns the number of final states visited.
// WRITE == true : also stores one record per final state at out[obase + k].
template <bool WRITE>
__device__ __forceinline__ uint32_t walk_start(const ScanParams &p, const uint8_t *__restrict__ buf,
                                               const int32_t *__restrict__ s_s0, uint32_t tpos,
                                               uint32_t lim_t, uint32_t rec_pos,
                                               unsigned long long obase)
{
    int32_t state = s_s0[buf[tpos]];                       // :41
    if (state < 0) return 0;                               // :43
    uint32_t c = 0;
    const int32_t colmask = (1 << p.width_bit) - 1;
    if (state < p.n_final) {                               // :44-47
        if (WRITE && obase + c < p.cap) p.out[obase + c] = make_uint2(rec_pos, (uint32_t)__ldg(&p.idmap[state]));
        c++;
    }
    uint32_t q = tpos + 1;
    while (q < lim_t) {                                    // :50
        const int32_t key = (state << 8) + buf[q];         // :52
        const int32_t row = key >> p.width_bit;            // :53
        const int32_t idx = __ldg(&p.r[row]) + (key & colmask);   // :54-55
        if (idx < 0 || idx >= p.ht_size) break;            // :56-57
        const int2 hv = __ldg(&p.htval[idx]);              // :59-61
        if (hv.x != row) break;
        state = hv.y;
        if (state < p.n_final) {                           // :67-70
            if (WRITE && obase + c < p.cap) p.out[obase + c] = make_uint2(rec_pos, (uint32_t)__ldg(&p.idmap[state]));
            c++;
        }
        q++;
    }
    return c;
}

// tile-relative walk bound of a start at tile-relative tpos
__device__ __forceinline__ uint32_t walk_limit(const ScanParams &p, uint32_t a0, uint32_t tpos)
{
    uint32_t lim_a = p.a_valid_end;
    if (p.use_ref_bound) {
        // reference tiles are 4096 bytes of global positions with a 512-byte halo
        const unsigned long long g = p.base_pos + (unsigned long long)(a0 + tpos - p.mis);
        const unsigned long long lim_g = (g & ~4095ull) + 4608ull;
        const unsigned long long lim2 = lim_g - p.base_pos + p.mis;
        if (lim2 < lim_a) lim_a = (uint32_t)lim2;
    }
    uint32_t lim_t = lim_a - a0;
    const uint32_t depth = tpos + p.max_pat_len;   // a walk reads at most max_pat_len bytes
    return lim_t < depth ? lim_t : depth;
}

__device__ __forceinline__ void issue_tile(const ScanParams &p, uint32_t tile, uint8_t *buf, uint64_t *bar)
{
    const uint32_t a0 = tile * (uint32_t)kTile;
    uint32_t nbytes = p.a_valid_end - a0;
    const uint32_t want = (uint32_t)kTile + p.halo;
    if (nbytes > want) nbytes = want;
    const uint32_t nb16 = nbytes & ~15u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, nb16);
    if (nb16) bulk_g2s(buf, p.in_al + a0, nb16, bar);
}

__global__ void __launch_bounds__(kThreads, 2) pfac_scan_kernel(const ScanParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *s_bitmap = reinterpret_cast<uint32_t *>(smem);
    int32_t *s_s0 = reinterpret_cast<int32_t *>(smem + 8192);
    uint64_t *s_mbar = reinterpret_cast<uint64_t *>(smem + 9216);                    // [2]
    uint32_t *s_tile = reinterpret_cast<uint32_t *>(smem + 9232);                    // [2]
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(smem + 9240);   // [1]
    uint32_t *s_wtot = reinterpret_cast<uint32_t *>(smem + 9248);                    // [2][kWarps]
    uint16_t *s_queue = reinterpret_cast<uint16_t *>(smem + kSmemFixed);
    uint8_t *s_in = smem + kSmemFixed + kWarps * kWarpRange * 2;
    const uint32_t stride = scan_buf_stride(p.halo);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < 2048; i += kThreads) s_bitmap[i] = __ldg(&p.bitmap2[i]);
    if (tid < 256) s_s0[tid] = __ldg(&p.s0[tid]);
    if (tid == 0) {
        mbar_init(&s_mbar[0], 1);
        mbar_init(&s_mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int b = 0; b < 2; b++) {
            const uint32_t t = atomicAdd(p.ticket, 1u);
            s_tile[b] = t;
            if (t < p.n_tiles) issue_tile(p, t, s_in + b * stride, &s_mbar[b]);
        }
    }
    __syncthreads();

    uint16_t *wq = s_queue + warp * kWarpRange;
    for (uint32_t it = 0;; it++) {
        const int b = it & 1;
        const uint32_t tile = s_tile[b];
        if (tile >= p.n_tiles) break;
        uint8_t *buf = s_in + b * stride;
        const uint32_t a0 = tile * (uint32_t)kTile;
        {   // wait for the bulk copy of this tile
            const uint32_t parity = (it >> 1) & 1u;
            unsigned spins = 0;
            while (!mbar_try_wait(&s_mbar[b], parity)) {
                if (++spins > kSpinLimit) { atomicExch(p.error_flag, 2u); break; }
            }
        }
        uint32_t avail = p.a_valid_end - a0;
        const uint32_t want = (uint32_t)kTile + p.halo;
        if (avail > want) avail = want;
        if (avail & 15u) {   // last bytes of the input: not a whole 16-byte block, copied by hand
            const uint32_t nb16 = avail & ~15u;
            if ((uint32_t)tid < (avail & 15u)) buf[nb16 + tid] = p.in_al[(size_t)a0 + nb16 + tid];
            __syncthreads();
        }
        const bool edge = (a0 < p.mis) || (a0 + (uint32_t)kTile > p.a_start_end);

        // ---- phase 1: filter + ordered compaction into the warp queue
        uint32_t nq = 0;
#pragma unroll
        for (int st = 0; st < kWarpRange / 512; st++) {
            const uint32_t off = (uint32_t)warp * kWarpRange + st * 512 + lane * 16;
            const uint4 v = *reinterpret_cast<const uint4 *>(buf + off);
            uint32_t nx = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane == 31) nx = *reinterpret_cast<const uint32_t *>(buf + off + 16);
            uint32_t mask = (p.debug & 4u) ? 0u : filter16(v, nx, s_bitmap);
            if (edge) {   // start positions are [mis, a_start_end) in aligned coordinates
                const uint32_t a = a0 + off;
                const uint32_t lo = p.mis > a ? p.mis - a : 0u;
                const uint32_t hi = p.a_start_end > a ? p.a_start_end - a : 0u;
                uint32_t keep = hi >= 16u ? 0xffffu : ((1u << hi) - 1u);
                keep &= lo >= 16u ? 0u : (0xffffu << lo);
                mask &= keep;
            }
            uint32_t incl = __popc(mask);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            uint32_t q = nq + incl - __popc(mask);
            nq += __shfl_sync(0xffffffffu, incl, 31);
            while (mask) {
                const uint32_t bit = __ffs(mask) - 1;
                wq[q++] = (uint16_t)(off + bit);
                mask &= mask - 1;
            }
        }
        __syncwarp();

        // ---- phase 2: walk the queue, count matches (lane owns a contiguous run of entries)
        const uint32_t per = (nq + 31u) >> 5;
        const uint32_t qb = lane * per;
        const uint32_t qe = (qb + per < nq) ? qb + per : nq;
        uint32_t csum = 0;
        if (!(p.debug & 1u))
        for (uint32_t e = qb; e < qe; e++) {
            const uint32_t tpos = wq[e];
            csum += walk_start<false>(p, buf, s_s0, tpos, walk_limit(p, a0, tpos), 0u, 0ull);
        }
        uint32_t incl = csum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        const uint32_t lane_off = incl - csum;
        if (lane == 31) s_wtot[b * kWarps + warp] = incl;
        __syncthreads();   // S1: every warp is done reading this tile for phase 2

        // ---- phase 3: tile total, look-back, ordered emit
        uint32_t wv = lane < kWarps ? s_wtot[b * kWarps + lane] : 0u;
        uint32_t winc = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        const uint32_t T = __shfl_sync(0xffffffffu, winc, 31);
        const uint32_t woff = __shfl_sync(0xffffffffu, winc - wv, warp);

        if (T == 0) {
            if (tid == 0) {   // nobody reads this buffer again: refill it before the look-back
                const uint32_t t = atomicAdd(p.ticket, 1u);
                s_tile[b] = t;
                if (t < p.n_tiles) issue_tile(p, t, buf, &s_mbar[b]);
            }
            if (warp == 0 && !(p.debug & 2u)) {
                const unsigned long long excl = tile_lookback(p.tile_state, tile, 0ull, lane, p.error_flag);
                if (lane == 0 && tile == p.n_tiles - 1) *p.count_out = excl;
            }
        } else {
            if (warp == 0) {
                const unsigned long long excl = tile_lookback(p.tile_state, tile, (unsigned long long)T, lane, p.error_flag);
                if (lane == 0) {
                    *s_base = excl;
                    if (tile == p.n_tiles - 1) *p.count_out = excl + T;
                }
            }
            __syncthreads();   // S2: tile base visible
            if (csum) {
                unsigned long long o = *s_base + woff + lane_off;
                for (uint32_t e = qb; e < qe; e++) {
                    const uint32_t tpos = wq[e];
                    o += walk_start<true>(p, buf, s_s0, tpos, walk_limit(p, a0, tpos), a0 + tpos - p.mis + p.pos_bias, o);
                }
            }
            __syncthreads();   // S3: buffer and queue free
            if (tid == 0) {
                const uint32_t t = atomicAdd(p.ticket, 1u);
                s_tile[b] = t;
                if (t < p.n_tiles) issue_tile(p, t, buf, &s_mbar[b]);
            }
        }
    }
}

#endif  // __CUDACC__

}  // namespace pfac
