/*
 * oracle/pfac_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement, in plain C, of the reference PHFPFAC algorithm
 * (mickeyjoe666/PHFPFAC, regex_GPU_PHF/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's
 * library.  The product (phfpfac_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_*.py)
 *   - against the reference's own table builder compiled from /root/reference
 *     (oracle/_ref/libphfpfac_ref.so, recipe in oracle/Makefile): identical
 *     state_num / s0Table / r / HT / val / HTSize / patternIdMap,
 *   - against the known-answer counts in the reference tree
 *     (experiment/xaarecord etc., tmp.dat) and the result-file md5s of
 *     SURVEY.md section 8(c).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * regex_GPU_PHF/).  Deliberately NOT restated (undefined behaviour in the
 * reference, defined away and documented in DESIGN.md):
 *   - start positions >= input_size in the last 4 KiB tile (master_kernel.cu:40
 *     compares a tile-local position with the global size),
 *   - device bytes past input_size read through a tile's 512-byte halo
 *     (master_kernel.cu:223 allocates, :359 copies only input_size bytes),
 *   - zero-length patterns (create_table_reorder.c:362 reads pat[0] of a
 *     0-byte malloc) and pattern files without a trailing newline
 *     (create_table_reorder.c:71-77 spins to the "length over 1024" exit).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CHAR_SET 256                 /* ctdef.h:12 */
#define REF_ROW_MAX 1048576          /* PHF/phf.c:7 */
#define REF_COL_MAX 4096             /* PHF/phf.c:8 */
#define REF_HASHTABLE_MAX (163840 * 20) /* PHF/phf.c:10 */
#define REF_PAGE_SIZE_C 4096         /* master_kernel.cu:9-10 */
#define REF_HALO_C 512               /* master_kernel.cu:11 (128 ints) */
#define REF_MAX_PATTERN_BUF 1024     /* create_table_reorder.c:55,74 */

enum {
    ORACLE_OK = 0,
    ORACLE_ERR_IO = -1,
    ORACLE_ERR_PATTERN_TOO_LONG = -2,   /* also: file does not end with '\n' */
    ORACLE_ERR_EMPTY_PATTERN = -3,
    ORACLE_ERR_WIDTH = -4,
    ORACLE_ERR_ROW_MAX = -5,
    ORACLE_ERR_HT_FULL = -6,
    ORACLE_ERR_NOMEM = -7,
    ORACLE_ERR_DENSE_LIMIT = -8,
    ORACLE_ERR_ARG = -9
};

/* ctdef.h:17-23 */
typedef struct {
    int pattern_id;
    int pattern_len;
    char *pat;
} opattern;

typedef struct {
    int state_num;      /* create_table_reorder.c:376 */
    int n_final;        /* create_table_reorder.c:239,247 */
    int max_len;        /* create_table_reorder.c:319-321 */
    int state_cap;
    int **pfac;         /* [state][256], -1 = no edge (create_table_reorder.c:310) */
    int *idmap;         /* final state -> 1-based line number (create_table_reorder.c:318) */
    /* PHF (phf.c:151) */
    int n_keys, max_key, max_row, max_offset, ht_size;
    int *r;             /* REF_ROW_MAX entries, like main.cc:73 */
    int *HT;            /* REF_HASHTABLE_MAX entries, main.cc:74 */
    int *val;           /* main.cc:75 */
} opart;

typedef struct {
    int n_parts;
    int n_patterns;
    int max_pat_len;    /* main.cc:59 */
    int width;
    opattern *patterns; /* sorted, 0-based here (reference keeps index 0 unused) */
    opart *parts;
} oracle_t;

/* ------------------------------------------------------------------ patterns */

/* create_table_reorder.c:21-45 */
static int comp_pat(const void *a, const void *b)
{
    const opattern *p1 = (const opattern *)a;
    const opattern *p2 = (const opattern *)b;
    int l1 = p1->pattern_len, l2 = p2->pattern_len;
    int m = l1 < l2 ? l1 : l2;
    int res = memcmp(p1->pat, p2->pat, (size_t)m);
    if (res == 0) {
        if (l1 > l2) return 1;
        if (l1 < l2) return -1;
        return 0;
    }
    return res;
}

/* create_table_reorder.c:53-122 (read_pattern): raw bytes, '\n' separated, ids are
 * 1-based line numbers (:100), no escape processing (:71 plain fgetc), buffer of
 * 1024 so a pattern holds at most 1022 bytes before the '\n' (:72-77).
 * The byte source here is a memory buffer; fgetc at EOF yields -1, stored as 0xFF,
 * until the 1024 limit trips -- restated as ORACLE_ERR_PATTERN_TOO_LONG. */
static int read_patterns_mem(const unsigned char *buf, size_t len, opattern **out, int *n_out)
{
    size_t cap = 1024, n = 0, i = 0;
    opattern *all = (opattern *)malloc(cap * sizeof(opattern));
    if (!all) return ORACLE_ERR_NOMEM;
    if (len == 0) { free(all); return ORACLE_ERR_PATTERN_TOO_LONG; }
    while (1) {
        size_t start = i;
        int str_len = 0;
        while (1) {
            int ch = (i < len) ? buf[i] : -1;
            i++;
            str_len++;
            if (str_len >= REF_MAX_PATTERN_BUF) { /* :74-77 */
                for (size_t k = 0; k < n; k++) free(all[k].pat);
                free(all);
                return ORACLE_ERR_PATTERN_TOO_LONG;
            }
            if (ch == '\n') { str_len -= 1; break; } /* :79-83 */
        }
        if (str_len == 0) {   /* reference UB, defined away */
            for (size_t k = 0; k < n; k++) free(all[k].pat);
            free(all);
            return ORACLE_ERR_EMPTY_PATTERN;
        }
        if (n == cap) {
            cap *= 2;
            opattern *t = (opattern *)realloc(all, cap * sizeof(opattern));
            if (!t) { free(all); return ORACLE_ERR_NOMEM; }
            all = t;
        }
        all[n].pattern_id = (int)n + 1;            /* :100 */
        all[n].pattern_len = str_len;              /* :101 */
        all[n].pat = (char *)malloc((size_t)str_len);
        memcpy(all[n].pat, buf + start, (size_t)str_len); /* :102-103 */
        n++;
        if (i >= len) break;                        /* :106-109 feof after peek */
    }
    qsort(all, n, sizeof(opattern), comp_pat);      /* :116 */
    *out = all;
    *n_out = (int)n;
    return ORACLE_OK;
}

/* ctdef.h:37-99 (fgetc_ext) over a memory buffer: merges a backslash and what follows into one
 * byte.  Returns the byte, REF_EOL for a raw newline (pattern separator), -1 at end of input.
 * fscanf("%3o") / fscanf("%2x") are restated as "up to 3 octal / 2 hex digits"; %2x skips leading
 * white space like scanf does.  (Not restated: scanf's optional sign and 0x prefix inside %x.) */
#define REF_EOL 0x10A   /* ctdef.h:13 */
static int fgetc_ext_mem(const unsigned char *buf, size_t len, size_t *pi)
{
    size_t i = *pi;
    int ch0 = i < len ? buf[i] : -1;
    i++;
    if (ch0 == '\\') {                                   /* :46 */
        int ch1 = i < len ? buf[i] : -1;
        i++;
        if (ch1 < 0) { *pi = i; return ch0; }             /* :50-52 feof */
        if (ch1 >= '0' && ch1 <= '9') {                   /* :55-59 isdigit, then %3o */
            int value = 0, nd = 0;
            i--;
            while (nd < 3 && i < len && buf[i] >= '0' && buf[i] <= '7') { value = value * 8 + (buf[i] - '0'); i++; nd++; }
            *pi = i;
            return (int)((char)value);
        }
        *pi = i;
        switch (ch1) {                                    /* :61-91 */
        case 'a': return '\a';
        case 'b': return '\b';
        case 't': return '\t';
        case 'n': return '\n';
        case 'v': return '\v';
        case 'f': return '\f';
        case 'r': return '\r';
        case '\'': case '\"': case '\\': return ch1;
        case 'x': {                                       /* :80-86 %2x */
            int value = 0, nd = 0;
            while (i < len && (buf[i] == ' ' || (buf[i] >= 9 && buf[i] <= 13))) i++;
            while (nd < 2 && i < len) {
                int c = buf[i], d;
                if (c >= '0' && c <= '9') d = c - '0';
                else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
                else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
                else break;
                value = value * 16 + d; i++; nd++;
            }
            *pi = i;
            return (int)((char)value);
        }
        default:                                          /* :87-90 not an escape */
            *pi = i - 1;
            return ch0;
        }
    }
    *pi = i;
    if (ch0 == '\n') return REF_EOL;                      /* :94-96 */
    return ch0;
}

/* create_table_reorder.c:131-185 (read_pattern_ext): like read_pattern but through fgetc_ext. */
static int read_patterns_ext_mem(const unsigned char *buf, size_t len, opattern **out, int *n_out)
{
    size_t cap = 1024, n = 0, i = 0;
    opattern *all = (opattern *)malloc(cap * sizeof(opattern));
    char str[REF_MAX_PATTERN_BUF];
    int rc = ORACLE_OK;
    if (!all) return ORACLE_ERR_NOMEM;
    if (len == 0) { free(all); return ORACLE_ERR_PATTERN_TOO_LONG; }
    while (1) {
        int str_len = 0;
        while (1) {
            int ch = fgetc_ext_mem(buf, len, &i);          /* :152 */
            if (str_len >= REF_MAX_PATTERN_BUF - 1) { rc = ORACLE_ERR_PATTERN_TOO_LONG; goto fail; }   /* :155-158 */
            str[str_len++] = (char)ch;
            if (ch == REF_EOL) { str_len -= 1; break; }    /* :160-164 */
        }
        if (str_len == 0) { rc = ORACLE_ERR_EMPTY_PATTERN; goto fail; }
        if (n == cap) {
            cap *= 2;
            opattern *t = (opattern *)realloc(all, cap * sizeof(opattern));
            if (!t) { rc = ORACLE_ERR_NOMEM; goto fail; }
            all = t;
        }
        all[n].pattern_id = (int)n + 1;                    /* :168 */
        all[n].pattern_len = str_len;
        all[n].pat = (char *)malloc((size_t)str_len);
        memcpy(all[n].pat, str, (size_t)str_len);
        n++;
        if (i >= len) break;                                /* :174-180 */
    }
    qsort(all, n, sizeof(opattern), comp_pat);              /* :183 */
    *out = all;
    *n_out = (int)n;
    return ORACLE_OK;
fail:
    for (size_t k = 0; k < n; k++) free(all[k].pat);
    free(all);
    return rc;
}

/* ------------------------------------------------------------------ PFAC trie */

static int part_grow(opart *p, int need)
{
    if (need <= p->state_cap) return 0;
    int cap = p->state_cap ? p->state_cap : 64;
    while (cap < need) cap *= 2;
    int **t = (int **)realloc(p->pfac, (size_t)cap * sizeof(int *));
    if (!t) return -1;
    p->pfac = t;
    for (int x = p->state_cap; x < cap; x++) {
        p->pfac[x] = (int *)malloc(CHAR_SET * sizeof(int));
        if (!p->pfac[x]) return -1;
        memset(p->pfac[x], 0xFF, CHAR_SET * sizeof(int)); /* :310 */
    }
    p->state_cap = cap;
    return 0;
}

/* create_table_reorder.c:277-378 (patternsToPFAC).  Final states 0..n-1 are the
 * index in the sorted partition (:366), state n is unused, the initial state is
 * n+1 (:288), interior states are numbered from n+2 in creation order (:292,331-333).
 * The reference pre-allocates INITIAL_PFAC_SIZE rows; rows here grow on demand
 * (same contents). */
static int patterns_to_pfac(const opattern *pats, int n, opart *p)
{
    int initial_state = n + 1;
    int state = initial_state;
    int state_count = initial_state + 1;
    p->n_final = n;
    p->max_len = 0;
    p->idmap = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    if (!p->idmap || part_grow(p, state_count + 1)) return ORACLE_ERR_NOMEM;
    for (int i = 0; i < n; i++) {
        const opattern *cur = &pats[i];
        int j, ch;
        p->idmap[i] = cur->pattern_id;                       /* :318 */
        if (cur->pattern_len > p->max_len) p->max_len = cur->pattern_len; /* :319 */
        for (j = 0; j < cur->pattern_len - 1; j++) {         /* :325 */
            ch = (unsigned char)cur->pat[j];
            if (p->pfac[state][ch] == -1) {                  /* :330 */
                p->pfac[state][ch] = state_count;
                state = state_count;
                state_count += 1;
                if (part_grow(p, state_count + 1)) return ORACLE_ERR_NOMEM;
            } else {
                state = p->pfac[state][ch];                  /* :357 */
            }
        }
        ch = (unsigned char)cur->pat[j];                     /* :362 */
        p->pfac[state][ch] = i;                              /* :366 */
        state = initial_state;                               /* :368 */
    }
    p->state_num = state_count;                              /* :376 */
    return ORACLE_OK;
}

/* ------------------------------------------------------------------ PHF (FFDM) */

typedef struct {      /* phf.c:15-19 */
    int RowNumber;
    int RowItemCnt;
    int *RowItemIdx;
} orow;

/* phf.c:126-139 (SortRows): the exact O(R^2) exchange sort; its tie order decides
 * the packing, so it is restated literally. */
static void sort_rows(int numRow, orow *Row)
{
    for (int i = 0; i < numRow - 1; i++)
        for (int j = i + 1; j < numRow; j++)
            if (Row[i].RowItemCnt < Row[j].RowItemCnt) {
                orow tmp = Row[i];
                Row[i] = Row[j];
                Row[j] = tmp;
            }
}

/* phf.c:151-291 (FFDM) with InitArrays (:62-77) and ReadKey (:90-117). */
static int ffdm(opart *p, int width)
{
    int ary_size = p->state_num;
    if (width > REF_COL_MAX) return ORACLE_ERR_WIDTH;        /* :161 */
    p->r = (int *)malloc(REF_ROW_MAX * sizeof(int));
    p->HT = (int *)malloc(REF_HASHTABLE_MAX * sizeof(int));
    p->val = (int *)malloc(REF_HASHTABLE_MAX * sizeof(int));
    if (!p->r || !p->HT || !p->val) return ORACLE_ERR_NOMEM;
    memset(p->r, 0xFF, REF_ROW_MAX * sizeof(int));           /* :67-69 */
    memset(p->HT, 0xFF, REF_HASHTABLE_MAX * sizeof(int));
    memset(p->val, 0xFF, REF_HASHTABLE_MAX * sizeof(int));

    /* ReadKey: keys ascending, key = state*256+ch, row = key/width, col = key%width */
    long long total = (long long)ary_size * CHAR_SET;
    int n_rows_alloc = (int)(total / width) + 2;
    if (n_rows_alloc > REF_ROW_MAX + 1) n_rows_alloc = REF_ROW_MAX + 1;
    orow *Row = (orow *)calloc((size_t)n_rows_alloc, sizeof(orow));
    if (!Row) return ORACLE_ERR_NOMEM;
    for (int x = 0; x < n_rows_alloc; x++) Row[x].RowNumber = x;   /* :72 */
    int KeyCount = 0, MaxKey = 0;
    for (long long key = 0; key < total; key++) {            /* :98 */
        if (p->pfac[key / CHAR_SET][key % CHAR_SET] < 0) continue;
        long long row = key / width;
        int col = (int)(key % width);
        if (row >= REF_ROW_MAX) { free(Row); return ORACLE_ERR_ROW_MAX; } /* :102 */
        orow *R = &Row[row];
        R->RowItemCnt += 1;
        R->RowItemIdx = (int *)realloc(R->RowItemIdx, (size_t)R->RowItemCnt * sizeof(int));
        R->RowItemIdx[R->RowItemCnt - 1] = col;              /* :107-109 */
        KeyCount++;
        if (key > MaxKey) MaxKey = (int)key;                 /* :111 */
    }
    int MaxRow = MaxKey / width + 1;                         /* :174 */
    sort_rows(MaxRow, Row);                                  /* :175 */

    int MaxOffset = 0;
    for (int ndx = 0; ndx < n_rows_alloc && Row[ndx].RowItemCnt > 0; ndx++) { /* :184 */
        int row = Row[ndx].RowNumber;
        int cnt = Row[ndx].RowItemCnt;
        int *cols = Row[ndx].RowItemIdx;
        int offset, i;
        for (offset = -cols[0]; offset < REF_HASHTABLE_MAX - width; offset++) { /* :188 */
            for (i = 0; i < cnt; i++)
                if (p->HT[offset + cols[i]] != -1) break;    /* :191 */
            if (i == cnt) {
                p->r[row] = offset;                          /* :197 */
                if (offset > MaxOffset) MaxOffset = offset;
                for (i = 0; i < cnt; i++) {
                    int col = cols[i];
                    long long key = (long long)row * width + col;   /* :205 */
                    p->HT[offset + col] = row;                      /* :211 */
                    p->val[offset + col] = p->pfac[key / CHAR_SET][key % CHAR_SET]; /* :216 */
                }
                break;
            }
        }
        if (offset == REF_HASHTABLE_MAX - width) {           /* :224 */
            free(Row);
            return ORACLE_ERR_HT_FULL;
        }
    }
    int HTSize = 0;
    for (int i = MaxOffset; i < MaxOffset + width; i++)      /* :232-236 */
        if (p->HT[i] >= 0 || p->val[i] >= 0) HTSize = i + 1;
    for (int x = 0; x < n_rows_alloc; x++) free(Row[x].RowItemIdx);
    free(Row);
    p->n_keys = KeyCount;
    p->max_key = MaxKey;
    p->max_row = MaxRow;
    p->max_offset = MaxOffset;
    p->ht_size = HTSize;
    return ORACLE_OK;
}

/* ------------------------------------------------------------------ build */

void oracle_free(oracle_t *o)
{
    if (!o) return;
    for (int g = 0; g < o->n_parts; g++) {
        opart *p = &o->parts[g];
        for (int x = 0; x < p->state_cap; x++) free(p->pfac[x]);
        free(p->pfac); free(p->idmap); free(p->r); free(p->HT); free(p->val);
    }
    for (int i = 0; i < o->n_patterns; i++) free(o->patterns[i].pat);
    free(o->patterns);
    free(o->parts);
    free(o);
}

/* create_table_reorder.c:201-251 (create_table_reorder) + :253-274 (divide_patterns)
 * + main.cc:120-126 (one FFDM per partition).  The reference always makes
 * n_parts = 4*streamnum partitions (GPU_S = 4, :207,217); n_parts is a parameter here
 * so tests can also build the single automaton the product scans with. */
static oracle_t *oracle_build_any(const unsigned char *buf, size_t len, int n_parts, int width, int escapes, int *err);
oracle_t *oracle_build_mem(const unsigned char *buf, size_t len, int n_parts, int width, int *err)
{
    return oracle_build_any(buf, len, n_parts, width, 0, err);
}
/* the same flow with read_pattern_ext as the front-end (unused by the reference's main, kept behind a flag) */
oracle_t *oracle_build_mem_ext(const unsigned char *buf, size_t len, int n_parts, int width, int *err)
{
    return oracle_build_any(buf, len, n_parts, width, 1, err);
}
static oracle_t *oracle_build_any(const unsigned char *buf, size_t len, int n_parts, int width, int escapes, int *err)
{
    int e = ORACLE_OK;
    oracle_t *o = (oracle_t *)calloc(1, sizeof(oracle_t));
    if (!o) { if (err) *err = ORACLE_ERR_NOMEM; return NULL; }
    if (n_parts < 1 || width < 1) { e = ORACLE_ERR_ARG; goto fail; }
    o->width = width;
    e = escapes ? read_patterns_ext_mem(buf, len, &o->patterns, &o->n_patterns)
                : read_patterns_mem(buf, len, &o->patterns, &o->n_patterns);
    if (e) goto fail;
    o->n_parts = n_parts;
    o->parts = (opart *)calloc((size_t)n_parts, sizeof(opart));
    if (!o->parts) { e = ORACLE_ERR_NOMEM; goto fail; }
    {
        int k = o->n_patterns / n_parts;                 /* :220 */
        int l = k + o->n_patterns % n_parts;             /* :222 */
        for (int g = 0; g < n_parts; g++) {
            int cnt = (g == n_parts - 1) ? l : k;        /* :260-272 */
            e = patterns_to_pfac(o->patterns + (size_t)g * k, cnt, &o->parts[g]);
            if (e) goto fail;
            if (o->parts[g].max_len > o->max_pat_len) o->max_pat_len = o->parts[g].max_len; /* :238 */
            e = ffdm(&o->parts[g], width);               /* main.cc:125 */
            if (e) goto fail;
        }
    }
    if (err) *err = ORACLE_OK;
    return o;
fail:
    if (err) *err = e;
    oracle_free(o);
    return NULL;
}

static unsigned char *read_file(const char *path, size_t *len)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    rewind(f);
    unsigned char *b = (unsigned char *)malloc((size_t)sz + 1);
    if (b && fread(b, 1, (size_t)sz, f) != (size_t)sz) { free(b); b = NULL; }
    fclose(f);
    *len = (size_t)sz;
    return b;
}

oracle_t *oracle_build_file(const char *pattern_file, int n_parts, int width, int *err)
{
    size_t len;
    unsigned char *b = read_file(pattern_file, &len);
    if (!b) { if (err) *err = ORACLE_ERR_IO; return NULL; }
    oracle_t *o = oracle_build_mem(b, len, n_parts, width, err);
    free(b);
    return o;
}

int oracle_n_parts(const oracle_t *o) { return o->n_parts; }
int oracle_n_patterns(const oracle_t *o) { return o->n_patterns; }
int oracle_max_pat_len(const oracle_t *o) { return o->max_pat_len; }

/* info[0..8] = state_num, n_final, max_len, ht_size, max_row, n_keys, max_key, max_offset, width */
void oracle_part_info(const oracle_t *o, int g, int *info)
{
    const opart *p = &o->parts[g];
    info[0] = p->state_num; info[1] = p->n_final; info[2] = p->max_len; info[3] = p->ht_size;
    info[4] = p->max_row;   info[5] = p->n_keys;  info[6] = p->max_key; info[7] = p->max_offset;
    info[8] = o->width;
}
const int *oracle_part_r(const oracle_t *o, int g) { return o->parts[g].r; }
const int *oracle_part_HT(const oracle_t *o, int g) { return o->parts[g].HT; }
const int *oracle_part_val(const oracle_t *o, int g) { return o->parts[g].val; }
const int *oracle_part_idmap(const oracle_t *o, int g) { return o->parts[g].idmap; }
/* main.cc:200: s0Table = PFAC[final_state_num + 1] */
const int *oracle_part_s0(const oracle_t *o, int g) { return o->parts[g].pfac[o->parts[g].n_final + 1]; }
/* dense PFAC row, for trie-level checks */
const int *oracle_part_pfac_row(const oracle_t *o, int g, int state) { return o->parts[g].pfac[state]; }
/* number of r entries the device receives: master_kernel.cu:221,293 */
int oracle_part_r_entries(const oracle_t *o, int g)
{
    return (int)(((long long)o->parts[g].state_num * CHAR_SET) / o->width + 1);
}

/* ------------------------------------------------------------------ scan */

typedef struct {
    const int *s0, *r, *HT, *val;
    int ht_size, width_bit, n_final;
} otab;

static int log2_shift(int width)
{
    int b;
    for (b = 0; (width >> b) != 1; b++) ;    /* master_kernel.cu:397-398 */
    return b;
}

/* One start position: master_kernel.cu:37-74 (SUBSEG_MATCH).  `bdy` is the absolute
 * end of the walk (exclusive).  Calls emit(state) for every final state visited, in
 * visiting order (slot k of d_match_result, :45-46,:68-69).  Returns the count. */
static inline int walk_start(const otab *t, const unsigned char *in, long long pos, long long bdy,
                             int *states_out, int cap)
{
    int matchi = 0;
    int state = t->s0[in[pos]];                       /* :41 */
    if (state < 0) return 0;                          /* :43 */
    if (state < t->n_final) { if (matchi < cap) states_out[matchi] = state; matchi++; } /* :44-47 */
    pos += 1;
    while (1) {
        if (pos >= bdy) break;                        /* :50 */
        int ch = in[pos];
        int key = (int)(((unsigned)state << 8) + (unsigned)ch);   /* :52 */
        int row = key >> t->width_bit;                /* :53 */
        int col = key & ((1 << t->width_bit) - 1);    /* :54 */
        int index = t->r[row] + col;                  /* :55 */
        if (index < 0 || index >= t->ht_size) state = -1;      /* :56-57 */
        else if (t->HT[index] == row) state = t->val[index];   /* :59-61 */
        else state = -1;
        if (state == -1) break;                       /* :66 */
        if (state < t->n_final) { if (matchi < cap) states_out[matchi] = state; matchi++; } /* :67-70 */
        pos += 1;
    }
    return matchi;
}

/* Absolute walk bound of a start position: the walk of a thread in tile gbid may read
 * tile-local bytes < 4608 (master_kernel.cu:141-144), the last tile up to input_size;
 * bytes at or past input_size are never valid input (see header). */
static inline long long walk_bound(long long pos, long long n)
{
    long long b = (pos / REF_PAGE_SIZE_C) * REF_PAGE_SIZE_C + REF_PAGE_SIZE_C + REF_HALO_C;
    return b < n ? b : n;
}

/* Faithful flow for small inputs: one dense [n][max_len_g] array per partition filled as
 * TraceTable_kernel does, then main.cc:304-324 merge and main.cc:341-349 emit order.
 * Outputs (pos, pattern id) records; returns the number of records (may exceed cap,
 * only cap are stored), or a negative error. */
long long oracle_scan_dense(const oracle_t *o, const unsigned char *input, long long n,
                            int64_t *pos_out, int32_t *id_out, long long cap)
{
    int mpl = o->max_pat_len;
    if (n < 0) return ORACLE_ERR_ARG;
    if (n == 0 || mpl == 0) return 0;
    /* main.cc:308,313 / master_kernel.cu:105: unsigned 32-bit index arithmetic */
    if ((unsigned long long)n * (unsigned long long)mpl >= (1ULL << 32) || n > INT_MAX)
        return ORACLE_ERR_DENSE_LIMIT;
    int32_t *agg = (int32_t *)malloc((size_t)n * mpl * sizeof(int32_t)); /* main.cc:304 */
    if (!agg) return ORACLE_ERR_NOMEM;
    memset(agg, 0xFF, (size_t)n * mpl * sizeof(int32_t));                 /* main.cc:305 */
    for (int g = 0; g < o->n_parts; g++) {
        const opart *p = &o->parts[g];
        int ml = p->max_len;
        if (ml == 0) continue;
        uint32_t *res = (uint32_t *)malloc((size_t)n * ml * sizeof(uint32_t));
        if (!res) { free(agg); return ORACLE_ERR_NOMEM; }
        memset(res, 0xFF, (size_t)n * ml * sizeof(uint32_t));   /* master_kernel.cu:236 */
        otab t = { p->pfac[p->n_final + 1], p->r, p->HT, p->val, p->ht_size,
                   log2_shift(o->width), p->n_final };
        int *tmp = (int *)malloc((size_t)(ml + 1) * sizeof(int));
        for (long long i = 0; i < n; i++) {
            int c = walk_start(&t, input, i, walk_bound(i, n), tmp, ml);
            if (c > ml) c = ml;   /* cannot happen: depth k reaches at most k finals */
            for (int k = 0; k < c; k++) res[(size_t)i * ml + k] = (uint32_t)tmp[k];
        }
        free(tmp);
        for (long long i = 0; i < n; i++) {                     /* main.cc:307-321 */
            size_t k = (size_t)i * mpl;
            while (k < (size_t)n * mpl && agg[k] != -1) k++;
            for (int j = 0; j < ml; j++) {
                uint32_t s = res[(size_t)i * ml + j];
                if (s != 0xFFFFFFFFu) agg[k++] = p->idmap[s];
                else break;
            }
        }
        free(res);
    }
    long long cnt = 0;
    for (long long i = 0; i < n; i++)                           /* main.cc:341-349 */
        for (int j = 0; j < mpl; j++) {
            int32_t v = agg[(size_t)i * mpl + j];
            if (v == -1) break;
            if (cnt < cap) { pos_out[cnt] = i; id_out[cnt] = v; }
            cnt++;
        }
    free(agg);
    return cnt;
}

/* Same record list without the dense arrays (for inputs beyond the reference's
 * 32-bit/dense limits): per position, partitions ascending, walk order within a
 * partition -- the order main.cc:304-349 produces. */
long long oracle_scan_compact(const oracle_t *o, const unsigned char *input, long long n,
                              int64_t *pos_out, int32_t *id_out, long long cap)
{
    if (n < 0) return ORACLE_ERR_ARG;
    int mpl = o->max_pat_len;
    if (n == 0 || mpl == 0) return 0;
    otab *tabs = (otab *)malloc((size_t)o->n_parts * sizeof(otab));
    int *tmp = (int *)malloc((size_t)(mpl + 1) * sizeof(int));
    int wb = log2_shift(o->width);
    for (int g = 0; g < o->n_parts; g++) {
        const opart *p = &o->parts[g];
        otab t = { p->pfac[p->n_final + 1], p->r, p->HT, p->val, p->ht_size, wb, p->n_final };
        tabs[g] = t;
    }
    long long cnt = 0;
    for (long long i = 0; i < n; i++) {
        long long b = walk_bound(i, n);
        for (int g = 0; g < o->n_parts; g++) {
            if (o->parts[g].max_len == 0) continue;
            int c = walk_start(&tabs[g], input, i, b, tmp, mpl);
            for (int k = 0; k < c; k++) {
                if (cnt < cap) { pos_out[cnt] = i; id_out[cnt] = o->parts[g].idmap[tmp[k]]; }
                cnt++;
            }
        }
    }
    free(tmp); free(tabs);
    return cnt;
}

/* CPU traversal of ONE PHF automaton given as the canonical arrays (the thread_data
 * fields of main.cc:19-32), parallel over contiguous position ranges.  This is the
 * cpu_baseline / --impl reference leg of bench.py: the reference ships no CPU matcher
 * (main.cc:239 only ever calls GPU_TraceTable), so the baseline is this port of
 * SUBSEG_MATCH (master_kernel.cu:37-74) on host threads.
 * If pos_out == NULL only counts.  Records come out in (position, walk order).
 * ref_tile_bound != 0 applies the 4096+512 tile walk bound (master_kernel.cu:141-144). */
long long oracle_scan_tables_omp(const int *s0, const int *r, const int *HT, const int *val,
                                 int ht_size, int width, int n_final, const int *idmap,
                                 int max_pat_len, int ref_tile_bound,
                                 const unsigned char *input, long long n, int nthreads,
                                 int64_t *pos_out, int32_t *id_out, long long cap)
{
    if (n <= 0 || max_pat_len <= 0) return 0;
    otab t = { s0, r, HT, val, ht_size, log2_shift(width), n_final };
    if (nthreads < 1) nthreads = 1;
    long long *counts = (long long *)calloc((size_t)nthreads + 1, sizeof(long long));
    long long chunk = (n + nthreads - 1) / nthreads;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            long long acc = 0;
            for (int k = 0; k < nthreads; k++) { long long c = counts[k]; counts[k] = acc; acc += c; }
            counts[nthreads] = acc;
            if (!pos_out) break;
        }
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
#endif
        for (int k = 0; k < nthreads; k++) {
            long long lo = (long long)k * chunk, hi = lo + chunk;
            if (hi > n) hi = n;
            int *tmp = (int *)malloc((size_t)(max_pat_len + 1) * sizeof(int));
            long long c = 0, base = counts[k];
            for (long long i = lo; i < hi; i++) {
                long long b = ref_tile_bound ? walk_bound(i, n) : n;
                int m = walk_start(&t, input, i, b, tmp, max_pat_len);
                if (pass == 1)
                    for (int q = 0; q < m; q++) {
                        long long w = base + c + q;
                        if (w < cap) { pos_out[w] = i; id_out[w] = idmap[tmp[q]]; }
                    }
                c += m;
            }
            if (pass == 0) counts[k] = c;
            free(tmp);
        }
    }
    long long total = counts[nthreads];
    free(counts);
    return total;
}

/* The same scan in ONE pass over the input (the timed CPU legs of bench.py): every thread walks its
 * position range once, appending its records to a buffer of its own; the buffers are then copied,
 * in range order, into pos_out / id_out (up to cap records).  Returns the total number of matches. */
long long oracle_scan_tables_omp_1pass(const int *s0, const int *r, const int *HT, const int *val,
                                       int ht_size, int width, int n_final, const int *idmap,
                                       int max_pat_len, int ref_tile_bound,
                                       const unsigned char *input, long long n, int nthreads,
                                       int64_t *pos_out, int32_t *id_out, long long cap)
{
    if (n <= 0 || max_pat_len <= 0) return 0;
    otab t = { s0, r, HT, val, ht_size, log2_shift(width), n_final };
    if (nthreads < 1) nthreads = 1;
    long long *counts = (long long *)calloc((size_t)nthreads + 1, sizeof(long long));
    int64_t **lpos = (int64_t **)calloc((size_t)nthreads, sizeof(int64_t *));
    int32_t **lid = (int32_t **)calloc((size_t)nthreads, sizeof(int32_t *));
    long long chunk = (n + nthreads - 1) / nthreads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
#endif
    for (int k = 0; k < nthreads; k++) {
        long long lo = (long long)k * chunk, hi = lo + chunk;
        if (hi > n) hi = n;
        int *tmp = (int *)malloc((size_t)(max_pat_len + 1) * sizeof(int));
        long long c = 0, room = 4096;
        int64_t *bp = (int64_t *)malloc((size_t)room * sizeof(int64_t));
        int32_t *bi = (int32_t *)malloc((size_t)room * sizeof(int32_t));
        for (long long i = lo; i < hi; i++) {
            long long b = ref_tile_bound ? walk_bound(i, n) : n;
            int m = walk_start(&t, input, i, b, tmp, max_pat_len);
            if (c + m > room) {
                while (c + m > room) room *= 2;
                bp = (int64_t *)realloc(bp, (size_t)room * sizeof(int64_t));
                bi = (int32_t *)realloc(bi, (size_t)room * sizeof(int32_t));
            }
            for (int q = 0; q < m; q++) { bp[c + q] = i; bi[c + q] = idmap[tmp[q]]; }
            c += m;
        }
        counts[k] = c;
        lpos[k] = bp;
        lid[k] = bi;
        free(tmp);
    }
    long long acc = 0;
    for (int k = 0; k < nthreads; k++) { long long c = counts[k]; counts[k] = acc; acc += c; }
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
#endif
    for (int k = 0; k < nthreads; k++) {
        long long base = counts[k], c = (k + 1 < nthreads ? counts[k + 1] : acc) - base;
        if (pos_out && base < cap) {
            long long m = c < cap - base ? c : cap - base;
            memcpy(pos_out + base, lpos[k], (size_t)m * sizeof(int64_t));
            memcpy(id_out + base, lid[k], (size_t)m * sizeof(int32_t));
        }
        free(lpos[k]);
        free(lid[k]);
    }
    free(counts); free(lpos); free(lid);
    return acc;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ writer */

/* main.cc:335-350: "At position %4d, match pattern %d\n" into GPU_match_result.txt.
 * Positions are int in the reference; printed here as 64-bit with the same width rule. */
int oracle_write_result(const char *path, const int64_t *pos, const int32_t *ids, long long count)
{
    FILE *f = fopen(path, "w");
    if (!f) return ORACLE_ERR_IO;
    for (long long i = 0; i < count; i++)
        fprintf(f, "At position %4lld, match pattern %d\n", (long long)pos[i], ids[i]);
    fclose(f);
    return ORACLE_OK;
}

/* main.cc:45-352 end to end on the CPU: argv = pattern file, streamnum, width, input
 * file; input_size = filesize - 1 (main.cc:138).  Returns number of records or <0. */
long long oracle_run_cli(const char *pattern_file, int streamnum, int width,
                         const char *input_file, const char *out_path, int dense)
{
    int err = 0;
    oracle_t *o = oracle_build_file(pattern_file, 4 * streamnum, width, &err); /* :207,217 */
    if (!o) return err;
    size_t len;
    unsigned char *in = read_file(input_file, &len);
    if (!in) { oracle_free(o); return ORACLE_ERR_IO; }
    long long n = (long long)len - 1;                       /* main.cc:138 */
    if (n < 0) n = 0;
    long long cnt = dense ? oracle_scan_dense(o, in, n, NULL, NULL, 0)
                          : oracle_scan_compact(o, in, n, NULL, NULL, 0);
    if (cnt >= 0) {
        int64_t *pos = (int64_t *)malloc((size_t)(cnt + 1) * sizeof(int64_t));
        int32_t *ids = (int32_t *)malloc((size_t)(cnt + 1) * sizeof(int32_t));
        if (dense) oracle_scan_dense(o, in, n, pos, ids, cnt);
        else oracle_scan_compact(o, in, n, pos, ids, cnt);
        if (oracle_write_result(out_path, pos, ids, cnt)) cnt = ORACLE_ERR_IO;
        free(pos); free(ids);
    }
    free(in);
    oracle_free(o);
    return cnt;
}

#ifdef ORACLE_MAIN
int main(int argc, char **argv)
{
    if (argc < 5) {
        fprintf(stderr, "usage: %s <pattern file name> <streamnum> <PHF width> <input file name> [out]\n", argv[0]);
        return 255;
    }
    long long c = oracle_run_cli(argv[1], atoi(argv[2]), atoi(argv[3]), argv[4],
                                 argc > 5 ? argv[5] : "GPU_match_result.txt", 0);
    if (c < 0) { fprintf(stderr, "oracle error %lld\n", c); return 1; }
    printf("%lld records\n", c);
    return 0;
}
#endif
