/*
 * ref_cpu.h -- CPU oracle for the k-mer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the algorithms in /root/reference/dna.c for the
 * path this repo accelerates.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (libdnagpu, the host C code, the Python binding) never does.
 *
 * Parity status: PINNED by the reference's own examples (test.sql:46-119,
 * README.md:66-135) -- see tests/test_oracle_kat.py -- and, when
 * oracle/_ref/libdnaref.so can be built (oracle/Makefile target `ref`), by
 * running the reference's unmodified dna.c against it on random inputs.
 * The only unpinned part is the numeric VALUE of kmer_hash (PostgreSQL's
 * hash_any lives in PG core, which is not vendored; no reference test prints
 * a hash).  Hash values only place rows in buckets; no result depends on them.
 */
#ifndef REF_CPU_H
#define REF_CPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: the ereport(ERROR) sites on the path */
enum {
    REF_OK = 0,
    REF_ERR_DNA_CHAR = 1,       /* dna.c:125,166 */
    REF_ERR_DNA_EMPTY = 2,      /* dna.c:160-161 */
    REF_ERR_KMER_LEN = 3,       /* dna.c:401-402,466-467 */
    REF_ERR_KMER_CHAR = 4,      /* dna.c:415,473 */
    REF_ERR_KMER_EMPTY = 5,     /* dna.c:460-461 */
    REF_ERR_K_RANGE = 6,        /* dna.c:772-773 */
    REF_ERR_PREFIX_LEN = 7,     /* dna.c:854-856 */
    REF_ERR_QKMER_LEN = 8,      /* dna.c:1106-1108 */
    REF_ERR_QKMER_CHAR = 9,     /* dna.c:893-895 */
    REF_ERR_QKMER_EMPTY = 10,   /* dna.c:877-879 */
    REF_ERR_QKMER_TOOLONG = 11  /* dna.c:883-885 */
};
const char *ref_errmsg(int code);

/* dna codec (dna.c:114-171) */
uint64_t ref_dna_words(uint64_t n_bases);
int ref_encode_dna(const char *seq, uint64_t n_bases, uint64_t *words);
void ref_decode_dna(const uint64_t *words, uint64_t n_bases, char *out /* n+1 */);

/* kmer codec (dna.c:397-515) */
int ref_kmer_make(const char *seq, uint64_t *bits, int *length);
int ref_decode_kmer(uint64_t bits, int length, char *out /* length+1 */);

/* generate_kmers (dna.c:743-837): faithful per-k-mer decode -> string ->
 * kmer_make, with 64-bit indices; and the equivalent two-word window form. */
uint64_t ref_kmer_rows(uint64_t n_bases, int k);
int ref_generate_kmers(const uint64_t *words, uint64_t n_bases, int k,
                       uint64_t *out, uint64_t *n_out);
int ref_generate_kmers_window(const uint64_t *words, uint64_t n_bases, int k,
                              uint64_t *out, uint64_t *n_out);

/* kmer_eq / kmer_hash (dna.c:655-668, 722-735) */
int ref_kmer_eq(uint64_t a_bits, int a_len, uint64_t b_bits, int b_len);
uint32_t ref_kmer_hash(uint64_t bits);

/* starts_with (dna.c:842-866).  *err = REF_ERR_PREFIX_LEN when prefix longer.
 * prefix_len == 32 uses the full mask (the reference's shift-by-64 is UB). */
int ref_starts_with(uint64_t kmer_bits, int kmer_len, uint64_t prefix_bits,
                    int prefix_len, int *err);
/* Same, but reproducing what x86 does with the UB shift (mask 0 at len 32). */
int ref_starts_with_x86(uint64_t kmer_bits, int kmer_len, uint64_t prefix_bits,
                        int prefix_len, int *err);

/* qkmer (dna.c:876-900) and contains (dna.c:1064-1135) */
int ref_validate_qkmer(const char *pattern);
int ref_nucleotide_matches(char nucleotide, char iupac);
int ref_contains(const char *pattern, uint64_t kmer_bits, int kmer_len, int *err);

/* generate_kmers ... WHERE [kmer ^@ prefix] [AND qkmer @> kmer], sequence order.
 * prefix_len == 0 / pattern == NULL disable a predicate. */
int ref_filter_kmers(const uint64_t *words, uint64_t n_bases, int k,
                     uint64_t prefix_bits, int prefix_len, const char *pattern,
                     uint64_t *out, uint64_t *n_out);

/* GROUP BY kmer: a hash aggregate keyed by ref_kmer_hash/ref_kmer_eq. */
typedef struct ref_agg ref_agg;
ref_agg *ref_agg_new(uint64_t expected_keys);
void ref_agg_free(ref_agg *agg);
int ref_agg_add(ref_agg *agg, uint64_t kmer_bits, uint64_t times);
uint64_t ref_agg_groups(const ref_agg *agg);
/* total = sum(count), distinct = count(*), unique = count(*) FILTER (count=1) */
void ref_agg_stats(const ref_agg *agg, uint64_t *total, uint64_t *distinct,
                   uint64_t *unique);
/* groups sorted by kmer bits ascending; arrays of ref_agg_groups() entries */
void ref_agg_sorted(const ref_agg *agg, uint64_t *kmers, uint64_t *counts);
/* order-independent digests of the grouped result (for sizes where a sorted
 * list does not fit): sum and xor of mix(kmer)*count and of mix(kmer ^ count) */
void ref_agg_digest(const ref_agg *agg, uint64_t digest[4]);
void ref_pairs_digest(const uint64_t *kmers, const uint64_t *counts, uint64_t n,
                      uint64_t digest[4]);

/* Whole query: generate_kmers (+ optional WHERE) -> GROUP BY.  `faithful` = 1
 * uses the per-k-mer decode/re-encode path and per-row predicate calls, 0 the
 * window form.  reads: n_seqs values of bases_per_seq bases at stride_words. */
int ref_count_query(const uint64_t *words, uint64_t n_seqs, uint64_t bases_per_seq,
                    uint64_t stride_words, int k, uint64_t prefix_bits, int prefix_len,
                    const char *pattern, int faithful, ref_agg *agg);
/* The same with `threads` POSIX threads over disjoint base/read ranges, each
 * with a private aggregate, merged at the end (bench baseline only). */
int ref_count_query_mt(const uint64_t *words, uint64_t n_seqs, uint64_t bases_per_seq,
                       uint64_t stride_words, int k, uint64_t prefix_bits,
                       int prefix_len, const char *pattern, int faithful, int threads,
                       ref_agg *agg);

/* The same query for inputs whose grouped result does not fit in memory (3.1 Gbp, k = 31): rows grouped in
 * `passes` x P disjoint hash partitions, each aggregated by ref_agg; the three aggregates add and the
 * digest (same definition as ref_agg_digest) combines across partitions.  Memory ~ rows / passes * 8 B. */
int ref_count_query_big(const uint64_t *words, uint64_t n_seqs, uint64_t bases_per_seq,
                        uint64_t stride_words, int k, uint64_t prefix_bits, int prefix_len,
                        const char *pattern, int faithful, int passes, int threads,
                        uint64_t stats[3], uint64_t digest[4]);

/* synthetic inputs (include/dnagpu_synth.h) */
void ref_synth_seq(uint64_t seed, uint32_t repeat_every, uint64_t n_bases,
                   uint64_t first_word, uint64_t n_words, uint64_t *words);
void ref_synth_reads(uint64_t seed, uint32_t repeat_every, uint64_t first_read,
                     uint64_t n_reads, uint32_t bases_per_read, uint32_t stride_words,
                     uint64_t *words);

#ifdef __cplusplus
}
#endif
#endif
