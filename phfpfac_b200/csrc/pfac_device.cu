// Device side of the C ABI (include/pfac_b200.h): contexts, table upload, the device-resident
// scan and the host pipeline.  Replaces GPU_Malloc_Memory / GPU_TraceTable / GPU_Free_memory
// (reference master_kernel.cu:188-524).  There is no CPU fallback anywhere in this file: a
// missing GPU or any CUDA error is reported as PFAC_ERR_NO_DEVICE / PFAC_ERR_CUDA.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>

#include "pfac_internal.h"
#include "pfac_kernel.cuh"

using namespace pfac;

#define CU_TRY(expr)                                                                         \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return set_error(PFAC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                             cudaGetErrorString(e_), __FILE__, __LINE__);                    \
    } while (0)

namespace {

struct Ctrl {   // device-side control block, zeroed before every launch
    unsigned int ticket;
    unsigned int error_flag;
    unsigned long long count;
};

struct Stage {   // one pipeline stage of pfac_scan_host
    cudaStream_t stream = nullptr;
    uint8_t *d_in = nullptr;
    size_t d_in_cap = 0;
    pfac_match *d_out = nullptr;
    size_t d_out_cap = 0;   // records
    Ctrl *d_ctrl = nullptr;
    unsigned long long *d_tile_state = nullptr;
    size_t tile_cap = 0;
    Ctrl *h_ctrl = nullptr;   // pinned
    cudaEvent_t done = nullptr;
};

}  // namespace

struct pfac_ctx {
    int device = 0;
    int sm_count = 0;
    int blocks_per_sm = 0;
    int n_streams = 1;
    size_t chunk_bytes = 0;
    // tables (device)
    int32_t *d_r = nullptr, *d_idmap = nullptr, *d_s0 = nullptr;
    int2 *d_htval = nullptr;
    uint32_t *d_bitmap2 = nullptr;
    int32_t ht_size = 0, width_bit = 0, n_final = 0, max_pat_len = 0;
    uint32_t halo = 16;
    size_t smem_bytes = 0;
    // scan_device scratch (own stream use)
    cudaStream_t own_stream = nullptr;
    Ctrl *d_ctrl = nullptr;
    Ctrl *h_ctrl = nullptr;
    unsigned long long *d_tile_state = nullptr;
    size_t tile_cap = 0;
    std::vector<Stage> stages;
    uint64_t info[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t debug = 0;   // PFAC_DEBUG env (timing experiments only)
    std::mutex mu;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int ensure_tiles(unsigned long long **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return PFAC_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    size_t n = std::max<size_t>(need, 4096);
    CU_TRY(cudaMalloc(buf, n * sizeof(unsigned long long)));
    *cap = n;
    return PFAC_OK;
}

// Enqueue one scan on `stream`.  d_ctrl / tile_state belong to that stream's user.
int launch_scan(pfac_ctx *ctx, const void *d_in, uint64_t n_starts, uint64_t n_valid, uint64_t base_pos,
                uint32_t pos_bias, void *d_out, uint64_t cap, Ctrl *d_ctrl, unsigned long long **tile_state, size_t *tile_cap,
                cudaStream_t stream, uint64_t *tiles_out, uint64_t *ctas_out)
{
    if (n_valid < n_starts) return set_error(PFAC_ERR_ARG, "n_valid < n_starts");
    const uint64_t useful = n_starts + (uint64_t)(ctx->max_pat_len > 0 ? ctx->max_pat_len - 1 : 0);
    if (n_valid > useful) n_valid = useful;
    const uint32_t mis = (uint32_t)((uintptr_t)d_in & 15u);
    if (n_valid + mis + (uint64_t)kTile + 4096 >= (1ull << 32))
        return set_error(PFAC_ERR_LIMIT, "one scan call covers less than 4 GiB; split the input");
    CU_TRY(cudaMemsetAsync(d_ctrl, 0, sizeof(Ctrl), stream));
    if (tiles_out) *tiles_out = 0;
    if (ctas_out) *ctas_out = 0;
    if (n_starts == 0 || ctx->max_pat_len == 0) return PFAC_OK;
    ScanParams p;
    p.in_al = (const uint8_t *)d_in - mis;
    p.mis = mis;
    p.a_start_end = (uint32_t)(mis + n_starts);
    p.a_valid_end = (uint32_t)(mis + n_valid);
    p.n_tiles = (uint32_t)(((uint64_t)p.a_start_end + kTile - 1) / kTile);
    p.halo = ctx->halo;
    p.max_pat_len = (uint32_t)ctx->max_pat_len;
    p.use_ref_bound = ctx->max_pat_len > kRefHalo + 1;   // only then can the 4096+512 bound cut a walk
    p.base_pos = base_pos;
    p.pos_bias = pos_bias;
    p.r = ctx->d_r;
    p.htval = ctx->d_htval;
    p.idmap = ctx->d_idmap;
    p.s0 = ctx->d_s0;
    p.bitmap2 = ctx->d_bitmap2;
    p.ht_size = ctx->ht_size;
    p.width_bit = ctx->width_bit;
    p.n_final = ctx->n_final;
    p.out = (uint2 *)d_out;
    p.cap = cap;
    p.count_out = &d_ctrl->count;
    int e = ensure_tiles(tile_state, tile_cap, p.n_tiles);
    if (e) return e;
    p.tile_state = *tile_state;
    p.ticket = &d_ctrl->ticket;
    p.error_flag = &d_ctrl->error_flag;
    p.debug = ctx->debug;
    CU_TRY(cudaMemsetAsync(*tile_state, 0, (size_t)p.n_tiles * sizeof(unsigned long long), stream));
    const uint32_t grid = (uint32_t)std::min<uint64_t>(p.n_tiles, (uint64_t)ctx->sm_count * ctx->blocks_per_sm);
    pfac_scan_kernel<<<grid, kThreads, ctx->smem_bytes, stream>>>(p);
    CU_TRY(cudaGetLastError());
    if (tiles_out) *tiles_out = p.n_tiles;
    if (ctas_out) *ctas_out = grid;
    return PFAC_OK;
}

void free_stage(Stage &s)
{
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    if (s.d_ctrl) cudaFree(s.d_ctrl);
    if (s.d_tile_state) cudaFree(s.d_tile_state);
    if (s.h_ctrl) cudaFreeHost(s.h_ctrl);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Stage();
}

}  // namespace

extern "C" {

int pfac_device_count(int *count)
{
    if (!count) return set_error(PFAC_ERR_ARG, "null count");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return set_error(PFAC_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return PFAC_OK;
}

int pfac_ctx_create(int device, const pfac_tables *t, int part, int n_streams, size_t chunk_bytes, pfac_ctx **out)
{
    if (!out || !t || part < 0 || part >= (int)t->parts.size() || n_streams < 1 || n_streams > 64)
        return set_error(PFAC_ERR_ARG, "bad arguments to pfac_ctx_create");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return set_error(PFAC_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return set_error(PFAC_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard g(device);
    if (!g.ok) return set_error(PFAC_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    const Partition &P = t->parts[(size_t)part];
    std::unique_ptr<pfac_ctx> ctx(new pfac_ctx);
    ctx->device = device;
    ctx->n_streams = n_streams;
    ctx->chunk_bytes = chunk_bytes ? chunk_bytes : ((size_t)32 << 20);
    ctx->chunk_bytes = (ctx->chunk_bytes + kTile - 1) / kTile * kTile;
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_error(PFAC_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                         device, prop.major, prop.minor);
    ctx->sm_count = prop.multiProcessorCount;
    ctx->ht_size = P.ht_size;
    ctx->width_bit = width_bits(P.width);
    ctx->n_final = P.n_final;
    ctx->max_pat_len = P.max_len;
    ctx->halo = (uint32_t)std::max(16, ((P.max_len > 0 ? P.max_len - 1 : 0) + 15) / 16 * 16);
    ctx->smem_bytes = scan_smem_bytes(ctx->halo);
    if (const char *dbg = getenv("PFAC_DEBUG")) ctx->debug = (uint32_t)atoi(dbg);

    // canonical arrays -> device (r, {HT,val} interleaved, idmap, s0) + the 2-byte prefix bitmap
    const size_t n_r = std::max<size_t>(P.r.size(), 1), n_ht = std::max<size_t>((size_t)P.ht_size, 1);
    const size_t n_id = std::max<size_t>((size_t)P.n_final, 1);
    std::vector<int2> htval(n_ht, make_int2(-1, -1));
    for (int32_t i = 0; i < P.ht_size; i++) htval[(size_t)i] = make_int2(P.HT[(size_t)i], P.val[(size_t)i]);
    std::vector<uint32_t> bitmap(2048, 0);
    for (int b0 = 0; b0 < 256; b0++) {
        const int32_t s = P.s0.empty() ? -1 : P.s0[(size_t)b0];
        if (s < 0) continue;
        for (int b1 = 0; b1 < 256; b1++) {
            // a 1-byte pattern matches whatever follows; otherwise the walk must have a 2nd edge
            if (s < P.n_final || P.lookup(s, b1) >= 0) {
                const uint32_t win = (uint32_t)b0 | ((uint32_t)b1 << 8);
                bitmap[win >> 5] |= 1u << (win & 31u);
            }
        }
    }
    CU_TRY(cudaMalloc(&ctx->d_r, n_r * sizeof(int32_t)));
    CU_TRY(cudaMalloc(&ctx->d_htval, n_ht * sizeof(int2)));
    CU_TRY(cudaMalloc(&ctx->d_idmap, n_id * sizeof(int32_t)));
    CU_TRY(cudaMalloc(&ctx->d_s0, 256 * sizeof(int32_t)));
    CU_TRY(cudaMalloc(&ctx->d_bitmap2, 2048 * sizeof(uint32_t)));
    CU_TRY(cudaMemset(ctx->d_r, 0xFF, n_r * sizeof(int32_t)));
    CU_TRY(cudaMemset(ctx->d_idmap, 0, n_id * sizeof(int32_t)));
    if (!P.r.empty()) CU_TRY(cudaMemcpy(ctx->d_r, P.r.data(), P.r.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(ctx->d_htval, htval.data(), n_ht * sizeof(int2), cudaMemcpyHostToDevice));
    if (P.n_final) CU_TRY(cudaMemcpy(ctx->d_idmap, P.idmap.data(), (size_t)P.n_final * sizeof(int32_t), cudaMemcpyHostToDevice));
    std::vector<int32_t> s0(256, -1);
    if (!P.s0.empty()) s0 = P.s0;
    CU_TRY(cudaMemcpy(ctx->d_s0, s0.data(), 256 * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(ctx->d_bitmap2, bitmap.data(), 2048 * sizeof(uint32_t), cudaMemcpyHostToDevice));

    CU_TRY(cudaFuncSetAttribute(pfac_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_bytes));
    int bps = 0;
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, pfac_scan_kernel, kThreads, ctx->smem_bytes));
    if (bps < 1) return set_error(PFAC_ERR_CUDA, "scan kernel does not fit on an SM (smem %zu B)", ctx->smem_bytes);
    ctx->blocks_per_sm = bps;

    CU_TRY(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    CU_TRY(cudaMalloc(&ctx->d_ctrl, sizeof(Ctrl)));
    CU_TRY(cudaHostAlloc(&ctx->h_ctrl, sizeof(Ctrl), cudaHostAllocPortable));
    *out = ctx.release();
    return PFAC_OK;
}

void pfac_ctx_destroy(pfac_ctx *ctx)
{
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    for (auto &s : ctx->stages) free_stage(s);
    if (ctx->d_r) cudaFree(ctx->d_r);
    if (ctx->d_htval) cudaFree(ctx->d_htval);
    if (ctx->d_idmap) cudaFree(ctx->d_idmap);
    if (ctx->d_s0) cudaFree(ctx->d_s0);
    if (ctx->d_bitmap2) cudaFree(ctx->d_bitmap2);
    if (ctx->d_ctrl) cudaFree(ctx->d_ctrl);
    if (ctx->h_ctrl) cudaFreeHost(ctx->h_ctrl);
    if (ctx->d_tile_state) cudaFree(ctx->d_tile_state);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int pfac_ctx_device(const pfac_ctx *ctx) { return ctx ? ctx->device : -1; }

int pfac_scan_device(pfac_ctx *ctx, const void *d_in, uint64_t n_starts, uint64_t n_valid, uint64_t base_pos,
                     void *d_out, uint64_t cap, void *d_count, void *stream)
{
    if (!ctx || (!d_in && n_starts) || (!d_out && cap) || !d_count) return set_error(PFAC_ERR_ARG, "bad arguments to pfac_scan_device");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->own_stream;
    uint64_t tiles = 0, ctas = 0;
    int e = launch_scan(ctx, d_in, n_starts, n_valid, base_pos, 0u, d_out, cap, ctx->d_ctrl, &ctx->d_tile_state,
                        &ctx->tile_cap, st, &tiles, &ctas);
    if (e) return e;
    CU_TRY(cudaMemcpyAsync(d_count, &ctx->d_ctrl->count, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    ctx->info[0] = tiles ? 1 : 0;
    ctx->info[1] = tiles;
    ctx->info[2] = ctas;
    ctx->info[3] = ctx->smem_bytes;
    ctx->info[4] = ctx->info[5] = 0;
    ctx->info[6] = 1;
    return PFAC_OK;
}

int pfac_scan_device_sync(pfac_ctx *ctx, const void *d_in, uint64_t n_starts, uint64_t n_valid, uint64_t base_pos,
                          void *d_out, uint64_t cap, uint64_t *count, void *stream)
{
    if (!ctx || !count) return set_error(PFAC_ERR_ARG, "bad arguments to pfac_scan_device_sync");
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->own_stream;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        DeviceGuard g(ctx->device);
        uint64_t tiles = 0, ctas = 0;
        int e = launch_scan(ctx, d_in, n_starts, n_valid, base_pos, 0u, d_out, cap, ctx->d_ctrl, &ctx->d_tile_state,
                            &ctx->tile_cap, st, &tiles, &ctas);
        if (e) return e;
        CU_TRY(cudaMemcpyAsync(ctx->h_ctrl, ctx->d_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        ctx->info[0] = tiles ? 1 : 0;
        ctx->info[1] = tiles;
        ctx->info[2] = ctas;
        ctx->info[3] = ctx->smem_bytes;
        ctx->info[4] = ctx->info[5] = 0;
        ctx->info[6] = 1;
        if (ctx->h_ctrl->error_flag)
            return set_error(PFAC_ERR_INTERNAL, "device watchdog tripped (code %u)", ctx->h_ctrl->error_flag);
        *count = ctx->h_ctrl->count;
    }
    if (*count > cap) return set_error(PFAC_ERR_OUTPUT_FULL, "%llu matches exceed the capacity of %llu records",
                                       (unsigned long long)*count, (unsigned long long)cap);
    return PFAC_OK;
}

int pfac_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return set_error(PFAC_ERR_ARG, "null ptr");
    CU_TRY(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));   // main.cc:147
    return PFAC_OK;
}

void pfac_host_free(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
}

int pfac_ctx_last_scan_info(const pfac_ctx *ctx, uint64_t info[8])
{
    if (!ctx || !info) return set_error(PFAC_ERR_ARG, "bad arguments");
    memcpy(info, ctx->info, sizeof ctx->info);
    return PFAC_OK;
}

// H2D + kernel + D2H pipeline over n_streams stages (replaces the synchronous
// cudaMemcpy / launch / cudaMemcpy sequence of GPU_TraceTable, master_kernel.cu:359-428).
int pfac_scan_host(pfac_ctx *ctx, const void *h_in, uint64_t n_starts, uint64_t n_valid, uint64_t base_pos,
                   pfac_match *h_out, uint64_t cap, uint64_t *count)
{
    if (!ctx || (!h_in && n_starts) || (!h_out && cap) || !count || n_valid < n_starts)
        return set_error(PFAC_ERR_ARG, "bad arguments to pfac_scan_host");
    if (n_valid >= (1ull << 32)) return set_error(PFAC_ERR_LIMIT, "pfac_scan_host takes less than 4 GiB per call");
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    *count = 0;
    const uint64_t halo = ctx->max_pat_len > 0 ? (uint64_t)ctx->max_pat_len - 1 : 0;
    const uint64_t chunk = ctx->chunk_bytes;
    const uint64_t n_chunks = (n_starts + chunk - 1) / chunk;
    const int S = (int)std::min<uint64_t>((uint64_t)ctx->n_streams, std::max<uint64_t>(n_chunks, 1));
    if ((int)ctx->stages.size() < S) ctx->stages.resize((size_t)S);
    for (int s = 0; s < S; s++) {
        Stage &st = ctx->stages[(size_t)s];
        if (!st.stream) {
            CU_TRY(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
            CU_TRY(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
            CU_TRY(cudaMalloc(&st.d_ctrl, sizeof(Ctrl)));
            CU_TRY(cudaHostAlloc(&st.h_ctrl, sizeof(Ctrl), cudaHostAllocPortable));
        }
        const size_t need_in = (size_t)(std::min<uint64_t>(chunk, n_starts) + halo + 64);
        if (st.d_in_cap < need_in) {
            if (st.d_in) cudaFree(st.d_in);
            st.d_in = nullptr;
            st.d_in_cap = 0;
            CU_TRY(cudaMalloc(&st.d_in, need_in));
            st.d_in_cap = need_in;
        }
        const size_t need_out = std::max<size_t>((size_t)(std::min<uint64_t>(chunk, n_starts) / 8), 4096);
        if (st.d_out_cap < need_out) {
            if (st.d_out) cudaFree(st.d_out);
            st.d_out = nullptr;
            st.d_out_cap = 0;
            CU_TRY(cudaMalloc(&st.d_out, need_out * sizeof(pfac_match)));
            st.d_out_cap = need_out;
        }
    }
    uint64_t h2d = 0, d2h = 0, launches = 0, tiles_total = 0, ctas_max = 0;
    auto enqueue = [&](uint64_t c) -> int {
        Stage &st = ctx->stages[(size_t)(c % (uint64_t)S)];
        const uint64_t off = c * chunk;
        const uint64_t ns = std::min<uint64_t>(chunk, n_starts - off);
        const uint64_t nv = std::min<uint64_t>(ns + halo, n_valid - off);
        CU_TRY(cudaMemcpyAsync(st.d_in, (const uint8_t *)h_in + off, (size_t)nv, cudaMemcpyHostToDevice, st.stream));
        h2d += nv;
        uint64_t tiles = 0, ctas = 0;
        int e = launch_scan(ctx, st.d_in, ns, nv, base_pos + off, (uint32_t)off, st.d_out, st.d_out_cap, st.d_ctrl,
                            &st.d_tile_state, &st.tile_cap, st.stream, &tiles, &ctas);
        if (e) return e;
        launches += tiles ? 1 : 0;
        tiles_total += tiles;
        ctas_max = std::max(ctas_max, ctas);
        CU_TRY(cudaMemcpyAsync(st.h_ctrl, st.d_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st.stream));
        d2h += sizeof(Ctrl);
        CU_TRY(cudaEventRecord(st.done, st.stream));
        return PFAC_OK;
    };
    int rc = PFAC_OK;
    for (uint64_t c = 0; c < std::min<uint64_t>((uint64_t)S, n_chunks); c++)
        if ((rc = enqueue(c)) != PFAC_OK) return rc;
    uint64_t total = 0;
    for (uint64_t c = 0; c < n_chunks; c++) {
        Stage &st = ctx->stages[(size_t)(c % (uint64_t)S)];
        CU_TRY(cudaEventSynchronize(st.done));
        if (st.h_ctrl->error_flag)
            return set_error(PFAC_ERR_INTERNAL, "device watchdog tripped (code %u)", st.h_ctrl->error_flag);
        uint64_t m = st.h_ctrl->count;
        if (m > st.d_out_cap) {
            // dense matches: grow this stage's record buffer and scan the sub-chunk again
            CU_TRY(cudaStreamSynchronize(st.stream));
            cudaFree(st.d_out);
            st.d_out = nullptr;
            st.d_out_cap = 0;
            CU_TRY(cudaMalloc(&st.d_out, (size_t)m * sizeof(pfac_match)));
            st.d_out_cap = (size_t)m;
            if ((rc = enqueue(c)) != PFAC_OK) return rc;
            CU_TRY(cudaEventSynchronize(st.done));
            m = st.h_ctrl->count;
        }
        // records already carry positions relative to h_in[0] (pos_bias = sub-chunk offset)
        const uint64_t room = total < cap ? cap - total : 0;
        const uint64_t ncopy = std::min<uint64_t>(m, room);
        if (ncopy) {
            CU_TRY(cudaMemcpyAsync(h_out + total, st.d_out, (size_t)ncopy * sizeof(pfac_match),
                                   cudaMemcpyDeviceToHost, st.stream));
            d2h += ncopy * sizeof(pfac_match);
        }
        total += m;
        if (c + (uint64_t)S < n_chunks)
            if ((rc = enqueue(c + (uint64_t)S)) != PFAC_OK) return rc;
    }
    for (int s = 0; s < S; s++) CU_TRY(cudaStreamSynchronize(ctx->stages[(size_t)s].stream));
    *count = total;
    ctx->info[0] = launches;
    ctx->info[1] = tiles_total;
    ctx->info[2] = ctas_max;
    ctx->info[3] = ctx->smem_bytes;
    ctx->info[4] = h2d;
    ctx->info[5] = d2h;
    ctx->info[6] = n_chunks;
    if (total > cap)
        return set_error(PFAC_ERR_OUTPUT_FULL, "%llu matches exceed the capacity of %llu records",
                         (unsigned long long)total, (unsigned long long)cap);
    return PFAC_OK;
}

}  // extern "C"
