/*
 * dnagpu_synth.h -- seeded synthetic generator for 2-bit packed `dna` payloads.
 *
 * Header-only, plain C99, usable from host C, host C++ and CUDA device code
 * (every function is `static inline` and, under nvcc, __host__ __device__).
 *
 * Why it exists: the reference's own generator (data/create_dna.py:27-34) is
 * unseeded (`random.choice`), so its files cannot be reproduced.  This one is
 * position-addressable: word j of a stream is a pure function of (seed, j), so
 * the C harness, the CPU oracle and an on-device kernel all produce the same
 * words, and any rank of a multi-GPU run can generate just its own range.
 *
 * Layout produced = the reference's `Dna.bit_sequence` (dna.c:42-47,114-123):
 * base i lives at bits 2*(i%32)..+1 of word i/32, A=00 T=01 C=10 G=11, the
 * unused tail bits of the last word are zero (dna.c:186 palloc0).
 *
 * Repeat planting: uniform random data has (almost) no repeated k-mers for
 * k >= 21, so every R-th block of `block_words` words is a verbatim copy of an
 * earlier, never-planted block.  R = 0 disables planting.
 */
#ifndef DNAGPU_SYNTH_H
#define DNAGPU_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define DNAGPU_HD __host__ __device__
#else
#define DNAGPU_HD
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Default planting parameters (recorded in every bench JSON line). */
#define DNAGPU_SYNTH_REPEAT_EVERY 8u   /* R: every 8th block is a copy      */
#define DNAGPU_SYNTH_BLOCK_WORDS 32u   /* L: 32 words = 1024 bases per block */

static inline DNAGPU_HD uint64_t dnagpu_splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Raw (un-planted) word j of stream `seed`: 32 uniform bases. */
static inline DNAGPU_HD uint64_t dnagpu_synth_raw_word(uint64_t seed, uint64_t j)
{
    return dnagpu_splitmix64(dnagpu_splitmix64(seed) + j);
}

/*
 * Word j of a planted stream.  Blocks are `block_words` words long; block b is
 * a copy iff R > 0, b >= R and b % R == R-1.  Its source block is
 *   R * (h1 % (b / R)) + (h2 % (R-1))
 * i.e. a block of an earlier group whose in-group offset is < R-1, which is
 * never itself a copy (no chains), so the value is O(1) to compute.
 */
static inline DNAGPU_HD uint64_t dnagpu_synth_word(uint64_t seed, uint32_t R,
                                                   uint32_t block_words, uint64_t j)
{
    uint64_t b = j / block_words;
    if (R >= 2 && b >= R && (b % R) == (uint64_t)(R - 1)) {
        uint64_t h = dnagpu_splitmix64(dnagpu_splitmix64(seed ^ 0xA5A5A5A5DEADBEEFull) + b);
        uint64_t groups = b / R; /* >= 1 */
        uint64_t src = (uint64_t)R * ((h >> 20) % groups) + ((h & 0xFFFFFull) % (uint64_t)(R - 1));
        j = src * block_words + (j % block_words);
    }
    return dnagpu_synth_raw_word(seed, j);
}

/*
 * Word j (0-based) of a single synthetic sequence of n_bases bases.
 * Adds the fixed "edge" windows config 5 needs: words 8,9 are all ones
 * (64 x 'G' -> contains the k=32 all-ones k-mer) and words 16,17 are zero
 * (64 x 'A'), when the sequence has at least 32 words.  Tail bits are zeroed;
 * words at or past ceil(n_bases/32) are zero (pad).
 */
static inline DNAGPU_HD uint64_t dnagpu_synth_seq_word(uint64_t seed, uint32_t R,
                                                       uint64_t n_bases, uint64_t j)
{
    uint64_t n_words = (n_bases + 31) / 32;
    uint64_t w;
    if (j >= n_words) return 0;
    w = dnagpu_synth_word(seed, R, DNAGPU_SYNTH_BLOCK_WORDS, j);
    if (n_words >= 32) {
        if (j == 8 || j == 9) w = ~(uint64_t)0;
        if (j == 16 || j == 17) w = 0;
    }
    if (j == n_words - 1 && (n_bases % 32) != 0)
        w &= (((uint64_t)1 << (2 * (n_bases % 32))) - 1);
    return w;
}

/*
 * Word t (0 <= t < stride_words) of read r in a fixed-stride batch of reads,
 * each `bases_per_read` long.  A read is one "block" for planting purposes, so
 * every R-th read is an exact copy of an earlier read.  Words past the read's
 * own ceil(bases/32) words (stride padding) are zero.
 */
static inline DNAGPU_HD uint64_t dnagpu_synth_read_word(uint64_t seed, uint32_t R,
                                                        uint32_t bases_per_read,
                                                        uint32_t stride_words,
                                                        uint64_t r, uint32_t t)
{
    uint32_t wpr = (bases_per_read + 31) / 32;
    uint64_t w;
    (void)stride_words;
    if (t >= wpr) return 0;
    w = dnagpu_synth_word(seed, R, wpr, r * (uint64_t)wpr + t);
    if (t == wpr - 1 && (bases_per_read % 32) != 0)
        w &= (((uint64_t)1 << (2 * (bases_per_read % 32))) - 1);
    return w;
}

#ifdef __cplusplus
}
#endif
#endif /* DNAGPU_SYNTH_H */
