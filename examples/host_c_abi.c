/* Minimal C host for the drop-in boundary (include/pfac_b200.h), the way INTEGRATION.md binds it
 * into the reference's main.cc: build the tables, scan a file on every GPU, write
 * GPU_match_result.txt.  Plain C on purpose: the header must not need C++ or CUDA headers.
 *
 *   gcc -std=c11 -Iinclude examples/host_c_abi.c -Lphfpfac_b200/_build -lpfac_b200 \
 *       -Wl,-rpath,$PWD/phfpfac_b200/_build -o host_c_abi
 *   ./host_c_abi <pattern file> <width> [<input file>]
 * Without an input file only the table half runs (no GPU needed). */
#include <stdio.h>
#include <stdlib.h>

#include "pfac_b200.h"

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <pattern file> <width> [<input file>]\n", argv[0]);
        return 255;
    }
    pfac_tables *tab = NULL;
    if (pfac_tables_build_file(argv[1], 1, atoi(argv[2]), &tab)) {   /* CreateTable + FFDM, main.cc:100-126 */
        fprintf(stderr, "%s\n", pfac_last_error());
        return 1;
    }
    int32_t info[12];
    pfac_tables_part_info(tab, 0, info);
    printf("abi %d: %d patterns, max length %d, %d states, %d final, hash table %d\n", pfac_abi_version(),
           pfac_tables_n_patterns(tab), pfac_tables_max_pat_len(tab), info[0], info[1], info[3]);
    if (argc < 4) {
        pfac_tables_destroy(tab);
        return 0;
    }

    FILE *f = fopen(argv[3], "rb");
    if (!f) {
        perror("Open input file failed.");
        return 1;
    }
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    rewind(f);
    uint64_t n = size > 0 ? (uint64_t)size - 1 : 0;   /* main.cc:138 */
    void *input = NULL;
    if (pfac_host_alloc(&input, n + 1)) {              /* main.cc:147 */
        fprintf(stderr, "%s\n", pfac_last_error());
        return 1;
    }
    if (fread(input, 1, n, f) != n) return 1;          /* main.cc:154 */
    fclose(f);

    int n_gpu = 0;
    if (pfac_device_count(&n_gpu) || n_gpu < 1) {      /* main.cc:50 */
        fprintf(stderr, "no CUDA device: %s\n", pfac_last_error());
        return 1;
    }
    pfac_job *job = NULL;
    uint64_t n_matches = 0;
    if (pfac_job_create(tab, NULL, n_gpu, 4, 0, &job) || pfac_job_run(job, input, n, &n_matches)) {
        fprintf(stderr, "%s\n", pfac_last_error());
        return 1;
    }
    void *w = NULL;
    if (pfac_write_begin("GPU_match_result.txt", &w)) return 1;       /* main.cc:335 */
    for (int i = 0; i < pfac_job_n_segments(job); i++) {
        uint64_t base, cnt;
        const pfac_match *rec;
        pfac_job_segment(job, i, &base, &rec, &cnt);
        if (pfac_write_records(w, base, rec, cnt)) return 1;          /* main.cc:341-349 */
    }
    if (pfac_write_end(w)) return 1;
    printf("%llu matches on %d GPU(s)\n", (unsigned long long)n_matches, n_gpu);
    pfac_job_destroy(job);
    pfac_host_free(input);
    pfac_tables_destroy(tab);
    return 0;
}
