/*
 * pfac_synth.h -- seeded synthetic pattern sets and texts of BASELINE.json's configs
 * (SURVEY.md section 8(d)).  Workload generation for tests and bench.py -- a tools library
 * (tools/_build/libpfac_synth.so), NOT part of the product ABI; the reference ships only fixed files
 * (regex_GPU_PHF/bytefile/, xaa..xad), no generator.
 */
#ifndef PFAC_SYNTH_H
#define PFAC_SYNTH_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
    PFAC_SYNTH_PAT_PRINTABLE = 0, /* unique, length uniform in [min,max], bytes 0x21..0x7E (configs 2, 4) */
    PFAC_SYNTH_PAT_SNORT = 1      /* unique Snort-like literals: 60% vocabulary tokens with shared
                                     prefixes, 40% binary; clipped log-normal length (config 3) */
};
enum {
    PFAC_SYNTH_TEXT_PRINTABLE = 0, /* bytes 0x20..0x7E, '\n' every <= 120 bytes (configs 2, 4) */
    PFAC_SYNTH_TEXT_HTTP = 1       /* HTTP-like lines from the same vocabulary, mixed 50/50 with
                                      uniform bytes (config 3) */
};

/* Writes `count` '\n'-terminated patterns (a pattern file image) into out[cap].
 * Returns the number of bytes needed (call with out = NULL to size), negative on error. */
long long pfac_synth_patterns(int kind, int count, uint64_t seed, int min_len, int max_len,
                              uint8_t *out, size_t cap);

/* Fills out[0, n).  If pattern_bytes != NULL one randomly chosen pattern is planted at a
 * jittered offset inside every 64 KiB block.  n_threads <= 0: all host threads.  The result
 * does not depend on n_threads. */
int pfac_synth_text(int kind, uint64_t seed, uint8_t *out, size_t n, const uint8_t *pattern_bytes,
                    size_t pattern_len, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
