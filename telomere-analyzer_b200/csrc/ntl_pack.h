/* ntl_pack.h -- host packer interface (see ntl_pack.cpp) */
#ifndef NTL_PACK_H
#define NTL_PACK_H
#include <stdint.h>
#include <functional>

/* Pack one read (ASCII, L >= 1) into 2-bit position words {lo, hi} (8 bytes per 32 positions, position p = bit
 * (p - 1) & 31 of word (p - 1) >> 5); dst holds n_words >= ceil(L / 32) records, the ones beyond the read are zeroed.
 * rc != 0: write the reverse complement.  Returns 0, or 1 if the read has a letter other than A/C/G/T (any case);
 * the caller then re-packs it with ntl_pack_read_4bit. */
int ntl_pack_read_2bit(const char *s, int64_t L, int rc, uint32_t *dst, int64_t n_words);
/* ntl_pack_read_2bit writes with non-temporal stores and does not fence: call this once after a batch of reads,
 * before another agent (DMA engine, another thread) is told that the words are in memory. */
void ntl_pack_fence(void);
/* 4-bit position words {A, C, G, T} (Biostrings code bits; gap letters have none), 16 bytes per 32 positions.
 * Returns 0, or -1 if a letter is outside the Biostrings DNA alphabet. */
int ntl_pack_read_4bit(const char *s, int64_t L, int rc, uint32_t *dst, int64_t n_words);
/* Biostrings nibble of a pattern letter (IUPAC, either case), -1 if not allowed. */
int ntl_pattern_nibble(char c);
/* fn(begin, end) over [0, n) in grains, on n_threads std::threads (the caller's thread included). */
void ntl_parallel_for(int64_t n, int n_threads, int64_t grain, const std::function<void(int64_t, int64_t)> &fn);
#endif
