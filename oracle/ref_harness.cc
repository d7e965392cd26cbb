/*
 * oracle/ref_harness.cc -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin C-ABI harness around the UNMODIFIED reference table builder, compiled from the
 * sources where they lie under /root/reference (never copied into this repo):
 *   - regex_GPU_PHF/CreateTable/create_PFAC_table_reorder.c  compiled as C by
 *     oracle/Makefile (gcc -std=gnu11 -include limits.h; g++ would trap on the
 *     missing `return` at create_table_reorder.c:251),
 *   - regex_GPU_PHF/PHF/phf.c  #included below (it only compiles as C++: it uses
 *     `RowStruct` without `struct`, phf.c:62).
 * Output: oracle/_ref/libphfpfac_ref.so (git-ignored, travels to the GPU box).
 * It is used to pin oracle/pfac_oracle.c and the product's table builder bit-for-bit.
 * The reference kernel (master_kernel.cu) cannot be built with CUDA 12 (legacy texture
 * references, master_kernel.cu:30-32) and the reference has no CPU matcher, so the scan
 * itself is restated in pfac_oracle.c.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <unistd.h>
#include <fcntl.h>

#define CHAR_SET 256
#include "PHF/phf.c"   /* -I/root/reference/regex_GPU_PHF ; defines ROW_MAX, HASHTABLE_MAX, FFDM */

extern "C" {
/* ctdef.h:17-23 */
struct ref_pattern_s { int pattern_id; int pattern_len; char *pat; };
/* create_PFAC_table_reorder.c:6 */
int create_PFAC_table_reorder(char *patternfilename, int *state_num, int *final_state_num,
                              int streamnum, int *max_pat_len_arr, int *max_pat_len,
                              int ***PFACs, int **patternIdMaps);
/* create_table_reorder.c:53, :277 */
ref_pattern_s *read_pattern(char *patternfilename, int *pattern_num, ref_pattern_s all_pattern[]);
/* create_table_reorder.c:131 */
void read_pattern_ext(char *patternfilename, int *pattern_num, ref_pattern_s all_pattern[]);
int **patternsToPFAC(ref_pattern_s patterns[], int pattern_num, int **PFAC, int *max_pat_length,
                     int *state_num, int patternIdMap[]);
extern int INITIAL_PFAC_SIZE;   /* create_table_reorder.c:10 */
extern int INITIAL_SIZE;        /* create_table_reorder.c:9  */
}

struct ref_tables {
    int n_parts, max_pat_len, width, pfac_rows;
    int *state_num, *final_num, *max_len_arr, *HTSize;
    int ***PFACs;
    int **idmaps, **r, **HT, **val;
};

/* the reference chatters on stdout (phf.c:262-282, create_table_reorder.c:213-226) */
struct quiet {
    int saved;
    quiet() {
        fflush(stdout);
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1);
        close(nul);
    }
    ~quiet() {
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

static ref_tables *alloc_tables(int n_parts, int width)
{
    ref_tables *t = (ref_tables *)calloc(1, sizeof(ref_tables));
    t->n_parts = n_parts;
    t->width = width;
    t->state_num = (int *)calloc(n_parts, sizeof(int));
    t->final_num = (int *)calloc(n_parts, sizeof(int));
    t->max_len_arr = (int *)calloc(n_parts, sizeof(int));   /* main.cc:57 calloc */
    t->HTSize = (int *)calloc(n_parts, sizeof(int));
    t->PFACs = (int ***)calloc(n_parts, sizeof(int **));
    t->idmaps = (int **)calloc(n_parts, sizeof(int *));
    t->r = (int **)calloc(n_parts, sizeof(int *));
    t->HT = (int **)calloc(n_parts, sizeof(int *));
    t->val = (int **)calloc(n_parts, sizeof(int *));
    for (int g = 0; g < n_parts; g++) {                     /* main.cc:72-76 */
        t->r[g] = (int *)malloc(ROW_MAX * sizeof(int));
        t->HT[g] = (int *)malloc(HASHTABLE_MAX * sizeof(int));
        t->val[g] = (int *)malloc(HASHTABLE_MAX * sizeof(int));
    }
    return t;
}

extern "C" {

/* The reference flow of main.cc:108 + main.cc:120-126: 4*streamnum partitions.
 * pfac_rows bounds the per-partition row pre-allocation (the reference default of
 * 4,000,000 rows x 1 KiB per partition is set through the non-const global, no edit). */
ref_tables *ref_build(const char *pattern_file, int streamnum, int width, int pfac_rows)
{
    quiet q;
    int n_parts = 4 * streamnum;   /* create_table_reorder.c:207,217 (GPU_S = 4) */
    ref_tables *t = alloc_tables(n_parts, width);
    INITIAL_PFAC_SIZE = pfac_rows;
    t->pfac_rows = pfac_rows;
    INITIAL_SIZE = 100000;
    create_PFAC_table_reorder((char *)pattern_file, t->state_num, t->final_num, streamnum,
                              t->max_len_arr, &t->max_pat_len, t->PFACs, t->idmaps);
    for (int g = 0; g < n_parts; g++)
        t->HTSize[g] = FFDM(t->PFACs[g], t->state_num[g], width, t->r[g], t->HT[g], t->val[g]);
    return t;
}

/* One automaton over ALL patterns: read_pattern + patternsToPFAC + FFDM called directly
 * (what the reference would build with a single partition). */
ref_tables *ref_build_single(const char *pattern_file, int width, int pfac_rows)
{
    quiet q;
    ref_tables *t = alloc_tables(1, width);
    INITIAL_PFAC_SIZE = pfac_rows;
    t->pfac_rows = pfac_rows;
    INITIAL_SIZE = 100000;
    int n = 0;
    ref_pattern_s *all = (ref_pattern_s *)malloc((size_t)INITIAL_SIZE * sizeof(ref_pattern_s));
    all = read_pattern((char *)pattern_file, &n, all);
    t->PFACs[0] = (int **)malloc((size_t)INITIAL_PFAC_SIZE * sizeof(int *));
    t->idmaps[0] = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    t->PFACs[0] = patternsToPFAC(&all[1], n, t->PFACs[0], &t->max_len_arr[0], &t->state_num[0],
                                 t->idmaps[0]);
    t->final_num[0] = n;
    t->max_pat_len = t->max_len_arr[0];
    t->HTSize[0] = FFDM(t->PFACs[0], t->state_num[0], width, t->r[0], t->HT[0], t->val[0]);
    return t;
}

/* The same with the reference's escape-processing reader (read_pattern_ext + fgetc_ext). */
ref_tables *ref_build_single_ext(const char *pattern_file, int width, int pfac_rows)
{
    quiet q;
    ref_tables *t = alloc_tables(1, width);
    INITIAL_PFAC_SIZE = pfac_rows;
    t->pfac_rows = pfac_rows;
    INITIAL_SIZE = 100000;
    int n = 0;
    ref_pattern_s *all = (ref_pattern_s *)malloc((size_t)INITIAL_SIZE * sizeof(ref_pattern_s));
    read_pattern_ext((char *)pattern_file, &n, all);
    t->PFACs[0] = (int **)malloc((size_t)INITIAL_PFAC_SIZE * sizeof(int *));
    t->idmaps[0] = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    t->PFACs[0] = patternsToPFAC(&all[1], n, t->PFACs[0], &t->max_len_arr[0], &t->state_num[0],
                                 t->idmaps[0]);
    t->final_num[0] = n;
    t->max_pat_len = t->max_len_arr[0];
    t->HTSize[0] = FFDM(t->PFACs[0], t->state_num[0], width, t->r[0], t->HT[0], t->val[0]);
    return t;
}

int ref_n_parts(const ref_tables *t) { return t->n_parts; }
int ref_max_pat_len(const ref_tables *t) { return t->max_pat_len; }
/* info[0..3] = state_num, final_state_num, max_pat_len_arr, HTSize */
void ref_part_info(const ref_tables *t, int g, int *info)
{
    info[0] = t->state_num[g]; info[1] = t->final_num[g];
    info[2] = t->max_len_arr[g]; info[3] = t->HTSize[g];
}
const int *ref_part_r(const ref_tables *t, int g) { return t->r[g]; }
const int *ref_part_HT(const ref_tables *t, int g) { return t->HT[g]; }
const int *ref_part_val(const ref_tables *t, int g) { return t->val[g]; }
const int *ref_part_idmap(const ref_tables *t, int g) { return t->idmaps[g]; }
const int *ref_part_pfac_row(const ref_tables *t, int g, int state) { return t->PFACs[g][state]; }
/* main.cc:200 */
const int *ref_part_s0(const ref_tables *t, int g) { return t->PFACs[g][t->final_num[g] + 1]; }
int ref_row_max(void) { return ROW_MAX; }
int ref_hashtable_max(void) { return HASHTABLE_MAX; }

/* the reference never frees its tables; rows added by its realloc-doubling path
 * (create_table_reorder.c:336-352) are not tracked here and leak */
void ref_free(ref_tables *t)
{
    for (int g = 0; g < t->n_parts; g++) {
        free(t->r[g]); free(t->HT[g]); free(t->val[g]);
        if (t->PFACs[g]) {
            int rows = t->pfac_rows > t->state_num[g] ? t->pfac_rows : t->state_num[g];
            if (t->state_num[g] < t->pfac_rows)   /* no doubling happened */
                for (int x = 0; x < rows; x++) free(t->PFACs[g][x]);
            free(t->PFACs[g]);
        }
        free(t->idmaps[g]);
    }
    free(t->r); free(t->HT); free(t->val);
    free(t->state_num); free(t->final_num); free(t->max_len_arr); free(t->HTSize);
    free(t->PFACs); free(t->idmaps);
    free(t);
}

}  /* extern "C" */
