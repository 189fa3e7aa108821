/*
 * dna_host.h -- host-side C mirror of the reference's interface for the k-mer path.
 *
 * Same value layouts as the extension (struct Dna dna.c:42-47, struct Kmer dna.c:61-65,
 * struct Qkmer dna.c:81-84, minus the varlena header PostgreSQL owns), same function names
 * and argument meaning, same error texts -- but every set-valued operation (generate_kmers,
 * the ^@ / @> scans, GROUP BY kmer) is executed by libdnagpu on a B200.  There is no CPU
 * implementation of those here: without a GPU dnah_open() fails.
 *
 * Scalar, per-value functions (dna_make, kmer_make, qkmer_make, the text output functions)
 * stay host code exactly as in the reference: they are O(length) glue that runs once per
 * literal, not the data-parallel path.
 */
#ifndef DNA_HOST_H
#define DNA_HOST_H

#include <stdint.h>

#include "../../include/dnagpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct Dna {         /* dna.c:42-47 */
    uint64_t length;         /* nucleotides */
    uint64_t bit_sequence[]; /* 2 bits per base, base i at bits 2*(i%32) of word i/32 */
} Dna;

typedef struct Kmer {        /* dna.c:61-65 */
    int32_t length;
    uint64_t bit_sequence;
} Kmer;

typedef struct Qkmer {       /* dna.c:81-84 */
    char sequence[33];       /* IUPAC pattern, NUL-terminated, at most 32 characters */
} Qkmer;

typedef struct KmerCount {   /* one row of SELECT kmer, count(*) ... GROUP BY kmer */
    Kmer kmer;
    uint64_t count;
} KmerCount;

/* error text of the last failed call on this thread (the reference's ereport message) */
const char *dnah_last_error(void);

/* ---- scalar glue (host), as in the reference ---- */
Dna *dna_make(const char *sequence);             /* dna.c:178-202; NULL + error on bad input */
char *dna_to_str(const Dna *dna);                /* dna.c:209-215; caller frees */
void dna_free(Dna *dna);
int kmer_make(const char *sequence, Kmer *out);  /* dna.c:487-515; 0 ok */
int kmer_to_str(const Kmer *kmer, char out[33]); /* dna.c:520-526 */
int qkmer_make(const char *sequence, Qkmer *out);/* dna.c:908-930 */
int kmer_eq_internal(const Kmer *a, const Kmer *b); /* dna.c:655-668 */

/* ---- the GPU-backed path ---- */
typedef struct dnah_session dnah_session; /* one backend = one GPU context */
int dnah_open(dnah_session **out, int device);
void dnah_close(dnah_session *s);

/* SELECT * FROM generate_kmers(dna, k)                                   (dna.c:743-837)
 * rows are malloc'ed; *n_rows of them; each has length == k. */
int generate_kmers(dnah_session *s, const Dna *dna, int k, Kmer **rows, uint64_t *n_rows);

/* ... WHERE kmer ^@ prefix AND pattern @> kmer (either may be NULL)      (dna.c:842-866, 1091-1135) */
int generate_kmers_where(dnah_session *s, const Dna *dna, int k, const Kmer *prefix, const Qkmer *pattern,
                         Kmer **rows, uint64_t *n_rows);

/* SELECT kmer, count(*) FROM generate_kmers(dna, k) [WHERE ...] GROUP BY kmer and the
 * total / distinct / unique aggregates over it                          (README.md:107-135)
 * rows may be NULL when only the aggregates are wanted. */
int count_kmers(dnah_session *s, const Dna *dna, int k, const Kmer *prefix, const Qkmer *pattern,
                KmerCount **rows, uint64_t *n_rows, dnagpu_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
