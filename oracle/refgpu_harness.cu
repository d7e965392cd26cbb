// TEST/BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry point around the reference's own GPU path -- GPU_Malloc_Memory / GPU_TraceTable /
// GPU_Free_memory of /root/reference/regex_GPU_PHF/master_kernel.cu, compiled from where it lies by
// oracle/Makefile (target refgpu).  That file does not build on CUDA 12 (legacy texture references,
// master_kernel.cu:30-32,302-320,518-520); the recipe streams it through sed, replacing the three
// tex1Dfetch calls by __ldg on the array pointers the kernel already receives (:94-95) and dropping
// the bind/unbind calls -- nothing else -- and keeps no copy of the source.
// Used by tests/refgpu_bench.py: the reference kernel on the same B200 as a baseline, and its dense
// result as one more parity check of the product's records.
#include <cuda_runtime.h>
#include <stdio.h>
#include <time.h>

struct thread_data {   // main.cc:19-32 == master_kernel.cu:15-28 (duplicated textually there too)
    unsigned char *input_string;
    int input_size;
    int state_num;
    int final_state_num;
    unsigned int *match_result;
    int HTSize;
    int width;
    int *s0Table;
    int max_pat_len;
    int *r;
    int *HT;
    int *val;
};

// main.cc:35-37
int GPU_Malloc_Memory(thread_data dataset, unsigned char **d_input_string, int **d_r, int **d_hash_table,
                      unsigned int **d_match_result, int **d_val_table, int **d_s0Table);
int GPU_TraceTable(thread_data dataset, cudaStream_t stream, unsigned char *d_input_string, int *d_r, int *d_hash_table,
                   unsigned int *d_match_result, int *d_val_table, int *d_s0Table);
int GPU_Free_memory(unsigned char **d_input_string, int **d_r, int **d_hash_table, unsigned int **d_match_result,
                    int **d_val_table, int **d_s0Table);

static double now_ms()
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec / 1e6;
}

// One pass of the reference's per-(GPU, stream) flow (main.cc:180-272) for a single partition on
// device 0.  match_result: max_pat_len * input_size u32 (pinned by the caller, main.cc:161).
// ms[0] = GPU_Malloc_Memory (allocations + 0xFF memset of the dense result), ms[1] = GPU_TraceTable
// (H2D of input and tables, kernel, D2H of the dense result), ms[2] = GPU_Free_memory.
extern "C" int refgpu_scan(unsigned char *input, int input_size, int state_num, int final_state_num, int HTSize, int width,
                           int *s0Table, int max_pat_len, int *r, int *HT, int *val, unsigned int *match_result, double *ms)
{
    if (cudaSetDevice(0) != cudaSuccess) return -1;
    cudaFree(0);
    thread_data d;
    d.input_string = input;
    d.input_size = input_size;
    d.state_num = state_num;
    d.final_state_num = final_state_num;
    d.match_result = match_result;
    d.HTSize = HTSize;
    d.width = width;
    d.s0Table = s0Table;
    d.max_pat_len = max_pat_len;
    d.r = r;
    d.HT = HT;
    d.val = val;
    unsigned char *d_in = nullptr;
    int *d_r = nullptr, *d_ht = nullptr, *d_val = nullptr, *d_s0 = nullptr;
    unsigned int *d_res = nullptr;
    double t0 = now_ms();
    GPU_Malloc_Memory(d, &d_in, &d_r, &d_ht, &d_res, &d_val, &d_s0);
    cudaDeviceSynchronize();
    double t1 = now_ms();
    GPU_TraceTable(d, 0, d_in, d_r, d_ht, d_res, d_val, d_s0);
    cudaDeviceSynchronize();
    double t2 = now_ms();
    GPU_Free_memory(&d_in, &d_r, &d_ht, &d_res, &d_val, &d_s0);
    double t3 = now_ms();
    ms[0] = t1 - t0;
    ms[1] = t2 - t1;
    ms[2] = t3 - t2;
    fflush(stdout);
    return 0;
}
