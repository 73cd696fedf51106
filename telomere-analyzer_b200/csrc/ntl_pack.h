/* ntl_pack.h -- host packer interface (see ntl_pack.cpp) */
#ifndef NTL_PACK_H
#define NTL_PACK_H
#include <stdint.h>
#include <functional>

/* Pack one read (ASCII, L >= 1) into 2-bit quads {lo[4], hi[4]}; dst holds ceil(((L>>5)+1)/4) * 8 words.
 * rc != 0: write the reverse complement.  Returns 0, or 1 if the read has a letter other than A/C/G/T (any case);
 * the caller then re-packs it with ntl_pack_read_4bit. */
int ntl_pack_read_2bit(const char *s, int64_t L, int rc, uint32_t *dst);
/* 4-bit quads {A[4], C[4], G[4], T[4]} (Biostrings code bits; gap letters have none); dst holds n_quads * 16 words.
 * Returns 0, or -1 if a letter is outside the Biostrings DNA alphabet. */
int ntl_pack_read_4bit(const char *s, int64_t L, int rc, uint32_t *dst);
/* Biostrings nibble of a pattern letter (IUPAC, either case), -1 if not allowed. */
int ntl_pattern_nibble(char c);
/* fn(begin, end) over [0, n) in grains, on n_threads std::threads (the caller's thread included). */
void ntl_parallel_for(int64_t n, int n_threads, int64_t grain, const std::function<void(int64_t, int64_t)> &fn);
#endif
