/*
 * ref_cpu.c -- CPU oracle for the k-mer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of /root/reference/dna.c for the path this repo
 * accelerates (see ref_cpu.h for the parity status).  Every function cites
 * the reference lines it follows.  Nothing here is linked into libdnagpu.
 *
 * Deliberate deviations, all forced by undefined behaviour in the reference
 * (SURVEY.md section 2.2):
 *   Q1  starts_with with a 32-base prefix shifts by 64 (dna.c:862): the oracle
 *       uses the full mask; ref_starts_with_x86 reproduces what x86 does.
 *   Q2  max_calls = length - k + 1 wraps when length < k - 1 (dna.c:781): the
 *       oracle returns zero rows when length < k.
 *   Q3  `int` indices overflow at 2^30 bases (dna.c:757,805-807): 64-bit here.
 */
#include "ref_cpu.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dnagpu_synth.h"

/* ---- errors ---------------------------------------------------------------- */
const char *ref_errmsg(int code)
{
    switch (code) {
    case REF_OK: return "ok";
    case REF_ERR_DNA_CHAR: return "Invalid character in DNA sequence";       /* dna.c:125,166 */
    case REF_ERR_DNA_EMPTY: return "DNA sequence cannot be empty";             /* dna.c:161 */
    case REF_ERR_KMER_LEN: return "K-mer length must be between 1 and 32 nucleotides"; /* dna.c:402 */
    case REF_ERR_KMER_CHAR: return "Invalid character in K-mer sequence";      /* dna.c:473 */
    case REF_ERR_KMER_EMPTY: return "K-mer sequence cannot be empty";          /* dna.c:461 */
    case REF_ERR_K_RANGE: return "Invalid k value: must be between 1 and 32";  /* dna.c:773 */
    case REF_ERR_PREFIX_LEN: return "Prefix length cannot exceed kmer length"; /* dna.c:855 */
    case REF_ERR_QKMER_LEN: return "Qkmer pattern and kmer lengths do not match"; /* dna.c:1107 */
    case REF_ERR_QKMER_CHAR: return "Invalid character in qkmer pattern";      /* dna.c:894 */
    case REF_ERR_QKMER_EMPTY: return "qkmer pattern cannot be empty";          /* dna.c:878 */
    case REF_ERR_QKMER_TOOLONG: return "Qkmer pattern length cannot exceed 32 characters"; /* dna.c:884 */
    }
    return "unknown";
}

/* ---- dna codec --------------------------------------------------------------- */
/* dna.c:179-181: (2*length + 63) / 64 words */
uint64_t ref_dna_words(uint64_t n_bases) { return (n_bases * 2 + 63) / 64; }

/* dna.c:159-171 (validate) then dna.c:114-128 (encode) into zeroed words
 * (dna.c:186 palloc0).  An empty sequence is an error (dna.c:160-161). */
int ref_encode_dna(const char *seq, uint64_t n_bases, uint64_t *words)
{
    uint64_t i;
    if (seq == NULL || n_bases == 0) return REF_ERR_DNA_EMPTY;
    for (i = 0; i < n_bases; i++) {
        char c = seq[i];
        if (c != 'A' && c != 'T' && c != 'C' && c != 'G') return REF_ERR_DNA_CHAR;
    }
    memset(words, 0, ref_dna_words(n_bases) * sizeof(uint64_t));
    for (i = 0; i < n_bases; i++) {
        uint64_t offset = (i * 2) % 64;
        uint64_t index = i / 32;
        switch (seq[i]) {
        case 'A': break;
        case 'T': words[index] |= ((uint64_t)0x1 << offset); break;
        case 'C': words[index] |= ((uint64_t)0x2 << offset); break;
        case 'G': words[index] |= ((uint64_t)0x3 << offset); break;
        }
    }
    return REF_OK;
}

/* dna.c:135-152 */
void ref_decode_dna(const uint64_t *words, uint64_t n_bases, char *out)
{
    static const char letters[4] = {'A', 'T', 'C', 'G'};
    uint64_t i;
    for (i = 0; i < n_bases; i++) {
        uint64_t offset = (i * 2) % 64;
        uint64_t index = i / 32;
        out[i] = letters[(words[index] >> offset) & 0x3];
    }
    out[n_bases] = '\0';
}

/* ---- kmer codec ---------------------------------------------------------------- */
/* kmer_make (dna.c:487-515) = strlen + validate_kmer_sequence (dna.c:457-479)
 * + encode_kmer (dna.c:397-420).  'X' encodes as 00 (dna.c:413). */
int ref_kmer_make(const char *seq, uint64_t *bits, int *length)
{
    size_t len, i;
    uint64_t b = 0;
    if (seq == NULL || *seq == '\0') return REF_ERR_KMER_EMPTY;
    len = strlen(seq);
    if (len > 32) return REF_ERR_KMER_LEN;
    for (i = 0; i < len; i++) {
        char c = seq[i];
        if (c != 'A' && c != 'T' && c != 'C' && c != 'G' && c != 'X') return REF_ERR_KMER_CHAR;
    }
    for (i = 0; i < len; i++) {
        int offset = (int)i * 2;
        switch (seq[i]) {
        case 'A': break;
        case 'T': b |= ((uint64_t)0x1 << offset); break;
        case 'C': b |= ((uint64_t)0x2 << offset); break;
        case 'G': b |= ((uint64_t)0x3 << offset); break;
        case 'X': break;
        }
    }
    *bits = b;
    *length = (int)len;
    return REF_OK;
}

/* dna.c:428-452 */
int ref_decode_kmer(uint64_t bits, int length, char *out)
{
    static const char letters[4] = {'A', 'T', 'C', 'G'};
    int i;
    if (length <= 0 || length > 32) return REF_ERR_KMER_LEN;
    for (i = 0; i < length; i++) out[i] = letters[(bits >> (i * 2)) & 0x3];
    out[length] = '\0';
    return REF_OK;
}

/* ---- generate_kmers -------------------------------------------------------------- */
/* dna.c:781 with Q2: zero rows when the sequence is shorter than k. */
uint64_t ref_kmer_rows(uint64_t n_bases, int k)
{
    if (k <= 0 || k > 32) return 0;
    return n_bases >= (uint64_t)k ? n_bases - (uint64_t)k + 1 : 0;
}

/* One row of generate_kmers exactly as dna.c:800-826 produces it: decode k
 * bases to a string, then kmer_make() re-validates and re-encodes it. */
static int ref_generate_one(const uint64_t *words, uint64_t current_index, int k, uint64_t *bits)
{
    char kmer_sequence[33];
    int i, length;
    for (i = 0; i < k; i++) {
        uint64_t nucleotide_index = current_index + (uint64_t)i;
        uint64_t bit_offset = (nucleotide_index * 2) % 64;
        uint64_t chunk_index = nucleotide_index / 32;
        uint64_t b = (words[chunk_index] >> bit_offset) & 0x3;
        switch (b) {
        case 0x0: kmer_sequence[i] = 'A'; break;
        case 0x1: kmer_sequence[i] = 'T'; break;
        case 0x2: kmer_sequence[i] = 'C'; break;
        case 0x3: kmer_sequence[i] = 'G'; break;
        }
    }
    kmer_sequence[k] = '\0';
    return ref_kmer_make(kmer_sequence, bits, &length);
}

/* dna.c:743-837 */
int ref_generate_kmers(const uint64_t *words, uint64_t n_bases, int k, uint64_t *out,
                       uint64_t *n_out)
{
    uint64_t rows, i;
    if (k <= 0 || k > 32) return REF_ERR_K_RANGE; /* dna.c:772-773 */
    rows = ref_kmer_rows(n_bases, k);
    for (i = 0; i < rows; i++) {
        int rc = ref_generate_one(words, i, k, &out[i]);
        if (rc != REF_OK) return rc;
    }
    *n_out = rows;
    return REF_OK;
}

/* The two-word window a k-mer occupies in the packed stream:
 * equal to ref_generate_one for every input (tests/test_oracle.py checks). */
static inline uint64_t ref_window(const uint64_t *words, uint64_t n_words, uint64_t i, int k)
{
    uint64_t q = i / 32;
    unsigned s = (unsigned)((i % 32) * 2);
    uint64_t x = words[q] >> s;
    if (s != 0 && q + 1 < n_words) x |= words[q + 1] << (64 - s);
    if (k < 32) x &= (((uint64_t)1 << (2 * k)) - 1);
    return x;
}

int ref_generate_kmers_window(const uint64_t *words, uint64_t n_bases, int k, uint64_t *out,
                              uint64_t *n_out)
{
    uint64_t rows, i, n_words = ref_dna_words(n_bases);
    if (k <= 0 || k > 32) return REF_ERR_K_RANGE;
    rows = ref_kmer_rows(n_bases, k);
    for (i = 0; i < rows; i++) out[i] = ref_window(words, n_words, i, k);
    *n_out = rows;
    return REF_OK;
}

/* ---- kmer_eq / kmer_hash ------------------------------------------------------------ */
/* dna.c:655-668 */
int ref_kmer_eq(uint64_t a_bits, int a_len, uint64_t b_bits, int b_len)
{
    if (a_len != b_len) return 0;
    if (a_bits != b_bits) return 0;
    return 1;
}

/* dna.c:722-735: hash_any(&bit_sequence, 8).  hash_any is PostgreSQL core
 * (src/common/hashfn.c, Bob Jenkins' lookup3; effective pin PostgreSQL 16, see
 * SURVEY.md 8(c4)); restated here for an aligned little-endian 8-byte key. */
#define REF_ROT(x, k) (((x) << (k)) | ((x) >> (32 - (k))))
uint32_t ref_kmer_hash(uint64_t bits)
{
    uint32_t a, b, c;
    a = b = c = 0x9e3779b9u + 8u + 3923095u;
    b += (uint32_t)(bits >> 32);
    a += (uint32_t)bits;
    c ^= b; c -= REF_ROT(b, 14);
    a ^= c; a -= REF_ROT(c, 11);
    b ^= a; b -= REF_ROT(a, 25);
    c ^= b; c -= REF_ROT(b, 16);
    a ^= c; a -= REF_ROT(c, 4);
    b ^= a; b -= REF_ROT(a, 14);
    c ^= b; c -= REF_ROT(b, 24);
    return c;
}

/* ---- starts_with ---------------------------------------------------------------------- */
/* dna.c:842-866 */
int ref_starts_with(uint64_t kmer_bits, int kmer_len, uint64_t prefix_bits, int prefix_len,
                    int *err)
{
    uint64_t mask;
    if (err) *err = REF_OK;
    if (prefix_len > kmer_len) { /* dna.c:854-856 */
        if (err) *err = REF_ERR_PREFIX_LEN;
        return 0;
    }
    mask = prefix_len >= 32 ? ~(uint64_t)0 : (((uint64_t)1 << (2 * prefix_len)) - 1); /* Q1 */
    return prefix_bits == (kmer_bits & mask);
}

int ref_starts_with_x86(uint64_t kmer_bits, int kmer_len, uint64_t prefix_bits, int prefix_len,
                        int *err)
{
    uint64_t mask;
    if (err) *err = REF_OK;
    if (prefix_len > kmer_len) {
        if (err) *err = REF_ERR_PREFIX_LEN;
        return 0;
    }
    /* x86 SHL masks the count to 6 bits: 1 << 64 == 1 << 0 == 1, mask == 0 */
    mask = (((uint64_t)1 << ((2 * prefix_len) & 63)) - 1);
    return prefix_bits == (kmer_bits & mask);
}

/* ---- qkmer / contains ------------------------------------------------------------------- */
/* dna.c:876-900 */
int ref_validate_qkmer(const char *pattern)
{
    const char *p;
    if (pattern == NULL || *pattern == '\0') return REF_ERR_QKMER_EMPTY;
    if (strlen(pattern) > 32) return REF_ERR_QKMER_TOOLONG;
    for (p = pattern; *p; p++) {
        switch (*p) {
        case 'A': case 'T': case 'C': case 'G': case 'U': case 'W': case 'S': case 'M':
        case 'K': case 'R': case 'Y': case 'B': case 'D': case 'H': case 'V': case 'N':
            break;
        default:
            return REF_ERR_QKMER_CHAR;
        }
    }
    return REF_OK;
}

/* dna.c:1064-1086 */
int ref_nucleotide_matches(char nucleotide, char iupac)
{
    switch (iupac) {
    case 'A': return nucleotide == 'A';
    case 'T': return nucleotide == 'T';
    case 'C': return nucleotide == 'C';
    case 'G': return nucleotide == 'G';
    case 'U': return nucleotide == 'U';
    case 'W': return nucleotide == 'A' || nucleotide == 'T';
    case 'S': return nucleotide == 'C' || nucleotide == 'G';
    case 'M': return nucleotide == 'A' || nucleotide == 'C';
    case 'K': return nucleotide == 'G' || nucleotide == 'T';
    case 'R': return nucleotide == 'A' || nucleotide == 'G';
    case 'Y': return nucleotide == 'C' || nucleotide == 'T';
    case 'B': return nucleotide == 'C' || nucleotide == 'G' || nucleotide == 'T';
    case 'D': return nucleotide == 'A' || nucleotide == 'G' || nucleotide == 'T';
    case 'H': return nucleotide == 'A' || nucleotide == 'C' || nucleotide == 'T';
    case 'V': return nucleotide == 'A' || nucleotide == 'C' || nucleotide == 'G';
    case 'N': return 1;
    default: return 0;
    }
}

/* dna.c:1091-1135 */
int ref_contains(const char *pattern, uint64_t kmer_bits, int kmer_len, int *err)
{
    static const char letters[4] = {'A', 'T', 'C', 'G'};
    int qkmer_length = (int)strlen(pattern), i;
    if (err) *err = REF_OK;
    if (qkmer_length != kmer_len) { /* dna.c:1106-1108 */
        if (err) *err = REF_ERR_QKMER_LEN;
        return 0;
    }
    for (i = 0; i < qkmer_length; i++) {
        char nucleotide = letters[(kmer_bits >> (i * 2)) & 0x3];
        if (!ref_nucleotide_matches(nucleotide, pattern[i])) return 0;
    }
    return 1;
}

/* generate_kmers(...) AS k WHERE k ^@ prefix AND pattern @> k  (test.sql:67,86) */
int ref_filter_kmers(const uint64_t *words, uint64_t n_bases, int k, uint64_t prefix_bits,
                     int prefix_len, const char *pattern, uint64_t *out, uint64_t *n_out)
{
    uint64_t rows, i, m = 0;
    int err;
    if (k <= 0 || k > 32) return REF_ERR_K_RANGE;
    if (pattern != NULL) {
        int rc = ref_validate_qkmer(pattern);
        if (rc != REF_OK) return rc;
    }
    rows = ref_kmer_rows(n_bases, k);
    /* the per-row ERRORs of dna.c:854-856 / 1106-1108 fire on the first row */
    if (rows > 0) {
        if (prefix_len > k) return REF_ERR_PREFIX_LEN;
        if (pattern != NULL && (int)strlen(pattern) != k) return REF_ERR_QKMER_LEN;
    }
    for (i = 0; i < rows; i++) {
        uint64_t bits;
        int rc = ref_generate_one(words, i, k, &bits);
        if (rc != REF_OK) return rc;
        if (prefix_len > 0 && !ref_starts_with(bits, k, prefix_bits, prefix_len, &err)) continue;
        if (pattern != NULL && !ref_contains(pattern, bits, k, &err)) continue;
        if (out) out[m] = bits;
        m++;
    }
    *n_out = m;
    return REF_OK;
}

/* ---- GROUP BY kmer ------------------------------------------------------------------------ */
/* A hash aggregate in the manner of PostgreSQL's HashAggregate: bucket chosen
 * by kmer_hash (dna.c:722-735), candidates confirmed by kmer_eq (dna.c:655-668),
 * transition = count(*) + 1.  k is the same for every row of a query, so the
 * length half of kmer_eq is constant; it is still compared. */
typedef struct ref_slot {
    uint64_t key;
    uint64_t count; /* 0 = empty */
} ref_slot;

struct ref_agg {
    ref_slot *slots;
    uint64_t cap; /* power of two */
    uint64_t groups;
    int k;
    /* after a multi-threaded query the groups live in n_shards disjoint sub-aggregates
     * (shard = hash >> 8 mod n_shards); every reader below walks them */
    int n_shards;
    struct ref_agg **shards;
};

static uint64_t ref_pow2_at_least(uint64_t x)
{
    uint64_t p = 16;
    while (p < x) p <<= 1;
    return p;
}

ref_agg *ref_agg_new(uint64_t expected_keys)
{
    ref_agg *agg = (ref_agg *)calloc(1, sizeof(*agg));
    if (!agg) return NULL;
    agg->cap = ref_pow2_at_least(expected_keys + expected_keys / 2 + 16);
    agg->slots = (ref_slot *)calloc(agg->cap, sizeof(ref_slot));
    agg->k = 0;
    if (!agg->slots) {
        free(agg);
        return NULL;
    }
    return agg;
}

void ref_agg_free(ref_agg *agg)
{
    int t;
    if (!agg) return;
    for (t = 0; t < agg->n_shards; t++) ref_agg_free(agg->shards[t]);
    free(agg->shards);
    free(agg->slots);
    free(agg);
}

static int ref_agg_grow(ref_agg *agg);

int ref_agg_add(ref_agg *agg, uint64_t kmer_bits, uint64_t times)
{
    uint64_t mask, i;
    if (times == 0) return 0;
    if (agg->n_shards)
        return ref_agg_add(agg->shards[(ref_kmer_hash(kmer_bits) >> 8) % (uint32_t)agg->n_shards],
                           kmer_bits, times);
    if ((agg->groups + 1) * 4 > agg->cap * 3) {
        if (ref_agg_grow(agg) != 0) return -1;
    }
    mask = agg->cap - 1;
    i = (uint64_t)ref_kmer_hash(kmer_bits) & mask;
    for (;;) {
        ref_slot *s = &agg->slots[i];
        if (s->count == 0) {
            s->key = kmer_bits;
            s->count = times;
            agg->groups++;
            return 0;
        }
        if (ref_kmer_eq(s->key, agg->k, kmer_bits, agg->k)) {
            s->count += times;
            return 0;
        }
        i = (i + 1) & mask;
    }
}

static int ref_agg_grow(ref_agg *agg)
{
    ref_slot *old = agg->slots;
    uint64_t old_cap = agg->cap, i;
    agg->cap = old_cap * 2;
    agg->slots = (ref_slot *)calloc(agg->cap, sizeof(ref_slot));
    if (!agg->slots) {
        agg->slots = old;
        agg->cap = old_cap;
        return -1;
    }
    agg->groups = 0;
    for (i = 0; i < old_cap; i++)
        if (old[i].count) ref_agg_add(agg, old[i].key, old[i].count);
    free(old);
    return 0;
}

uint64_t ref_agg_groups(const ref_agg *agg)
{
    uint64_t g = agg->groups;
    int t;
    for (t = 0; t < agg->n_shards; t++) g += agg->shards[t]->groups;
    return g;
}

/* README.md:122-130 / test.sql:107-115: sum(count), count(*),
 * count(*) FILTER (WHERE count = 1) over the grouped rows. */
void ref_agg_stats(const ref_agg *agg, uint64_t *total, uint64_t *distinct, uint64_t *unique)
{
    uint64_t t = 0, d = 0, u = 0, i;
    int sh;
    for (i = 0; i < agg->cap; i++) {
        uint64_t c = agg->slots[i].count;
        if (!c) continue;
        t += c;
        d += 1;
        u += (c == 1);
    }
    for (sh = 0; sh < agg->n_shards; sh++) {
        uint64_t t2, d2, u2;
        ref_agg_stats(agg->shards[sh], &t2, &d2, &u2);
        t += t2;
        d += d2;
        u += u2;
    }
    *total = t;
    *distinct = d;
    *unique = u;
}

typedef struct ref_pair {
    uint64_t key, count;
} ref_pair;

static int ref_pair_cmp(const void *a, const void *b)
{
    uint64_t x = ((const ref_pair *)a)->key, y = ((const ref_pair *)b)->key;
    return x < y ? -1 : (x > y ? 1 : 0);
}

void ref_agg_sorted(const ref_agg *agg, uint64_t *kmers, uint64_t *counts)
{
    uint64_t groups = ref_agg_groups(agg);
    ref_pair *p = (ref_pair *)malloc((groups ? groups : 1) * sizeof(ref_pair));
    uint64_t i, m = 0;
    int sh;
    for (sh = -1; sh < agg->n_shards; sh++) {
        const ref_agg *a = sh < 0 ? agg : agg->shards[sh];
        for (i = 0; i < a->cap; i++)
            if (a->slots[i].count) {
                p[m].key = a->slots[i].key;
                p[m].count = a->slots[i].count;
                m++;
            }
    }
    qsort(p, m, sizeof(ref_pair), ref_pair_cmp);
    for (i = 0; i < m; i++) {
        kmers[i] = p[i].key;
        counts[i] = p[i].count;
    }
    free(p);
}

static inline void ref_digest_step(uint64_t digest[4], uint64_t kmer, uint64_t count)
{
    uint64_t a = dnagpu_splitmix64(kmer) * count;
    uint64_t b = dnagpu_splitmix64(kmer ^ (count * 0x9E3779B97F4A7C15ull));
    digest[0] += a;
    digest[1] ^= a;
    digest[2] += b;
    digest[3] ^= b;
}

void ref_agg_digest(const ref_agg *agg, uint64_t digest[4])
{
    uint64_t i;
    int sh;
    digest[0] = digest[1] = digest[2] = digest[3] = 0;
    for (sh = -1; sh < agg->n_shards; sh++) {
        const ref_agg *a = sh < 0 ? agg : agg->shards[sh];
        for (i = 0; i < a->cap; i++)
            if (a->slots[i].count) ref_digest_step(digest, a->slots[i].key, a->slots[i].count);
    }
}

void ref_pairs_digest(const uint64_t *kmers, const uint64_t *counts, uint64_t n,
                      uint64_t digest[4])
{
    uint64_t i;
    digest[0] = digest[1] = digest[2] = digest[3] = 0;
    for (i = 0; i < n; i++) ref_digest_step(digest, kmers[i], counts[i]);
}

/* ---- whole query ------------------------------------------------------------------------------ */
/* SELECT k.kmer, count(*) FROM [reads r,] generate_kmers(seq, k) AS k(kmer)
 * [WHERE k.kmer ^@ prefix AND pattern @> k.kmer] GROUP BY k.kmer
 * (README.md:107-116, test.sql:95-104, table form test.sql:140-150). */
static int ref_count_range(const uint64_t *words, uint64_t seq_first, uint64_t seq_last,
                           uint64_t bases_per_seq, uint64_t stride_words, uint64_t row_first,
                           uint64_t row_last, int k, uint64_t prefix_bits, int prefix_len,
                           const char *pattern, int faithful, ref_agg *agg)
{
    uint64_t s, n_words = ref_dna_words(bases_per_seq);
    int err;
    agg->k = k;
    for (s = seq_first; s < seq_last; s++) {
        const uint64_t *w = words + s * stride_words;
        uint64_t rows = ref_kmer_rows(bases_per_seq, k), i;
        uint64_t lo = row_first, hi = row_last < rows ? row_last : rows;
        for (i = lo; i < hi; i++) {
            uint64_t bits;
            if (faithful) {
                int rc = ref_generate_one(w, i, k, &bits);
                if (rc != REF_OK) return rc;
                if (prefix_len > 0 && !ref_starts_with(bits, k, prefix_bits, prefix_len, &err))
                    continue;
                if (pattern != NULL && !ref_contains(pattern, bits, k, &err)) continue;
            } else {
                bits = ref_window(w, n_words, i, k);
                if (prefix_len > 0) {
                    uint64_t mask = prefix_len >= 32 ? ~(uint64_t)0
                                                     : (((uint64_t)1 << (2 * prefix_len)) - 1);
                    if ((bits & mask) != prefix_bits) continue;
                }
                if (pattern != NULL && !ref_contains(pattern, bits, k, &err)) continue;
            }
            if (ref_agg_add(agg, bits, 1) != 0) return -1;
        }
    }
    return REF_OK;
}

static int ref_query_check(uint64_t n_seqs, uint64_t bases_per_seq, int k, int prefix_len,
                           const char *pattern)
{
    if (k <= 0 || k > 32) return REF_ERR_K_RANGE;
    if (pattern != NULL) {
        int rc = ref_validate_qkmer(pattern);
        if (rc != REF_OK) return rc;
    }
    if (n_seqs > 0 && ref_kmer_rows(bases_per_seq, k) > 0) {
        if (prefix_len > k) return REF_ERR_PREFIX_LEN;
        if (pattern != NULL && (int)strlen(pattern) != k) return REF_ERR_QKMER_LEN;
    }
    return REF_OK;
}

int ref_count_query(const uint64_t *words, uint64_t n_seqs, uint64_t bases_per_seq,
                    uint64_t stride_words, int k, uint64_t prefix_bits, int prefix_len,
                    const char *pattern, int faithful, ref_agg *agg)
{
    int rc = ref_query_check(n_seqs, bases_per_seq, k, prefix_len, pattern);
    if (rc != REF_OK) return rc;
    return ref_count_range(words, 0, n_seqs, bases_per_seq, stride_words, 0, UINT64_MAX, k,
                           prefix_bits, prefix_len, pattern, faithful, agg);
}

typedef struct ref_mt_job {
    /* phase 2: shard `shard` of `n_jobs` collects its keys from every private aggregate */
    struct ref_mt_job *all;
    int n_jobs, shard;
    ref_agg *final;
    const uint64_t *words;
    uint64_t seq_first, seq_last, bases_per_seq, stride_words, row_first, row_last;
    int k, prefix_len, faithful, rc;
    uint64_t prefix_bits;
    const char *pattern;
    ref_agg *agg;
} ref_mt_job;

static void *ref_mt_main(void *arg)
{
    ref_mt_job *j = (ref_mt_job *)arg;
    j->rc = ref_count_range(j->words, j->seq_first, j->seq_last, j->bases_per_seq,
                            j->stride_words, j->row_first, j->row_last, j->k, j->prefix_bits,
                            j->prefix_len, j->pattern, j->faithful, j->agg);
    return NULL;
}

static void *ref_mt_merge(void *arg)
{
    ref_mt_job *j = (ref_mt_job *)arg;
    int t;
    for (t = 0; t < j->n_jobs; t++) {
        const ref_agg *a = j->all[t].agg;
        uint64_t i;
        for (i = 0; i < a->cap; i++)
            if (a->slots[i].count &&
                (int)((ref_kmer_hash(a->slots[i].key) >> 8) % (uint32_t)j->n_jobs) == j->shard)
                ref_agg_add(j->final, a->slots[i].key, a->slots[i].count);
    }
    return NULL;
}

int ref_count_query_mt(const uint64_t *words, uint64_t n_seqs, uint64_t bases_per_seq,
                       uint64_t stride_words, int k, uint64_t prefix_bits, int prefix_len,
                       const char *pattern, int faithful, int threads, ref_agg *agg)
{
    ref_mt_job *jobs;
    pthread_t *tids;
    uint64_t rows = ref_kmer_rows(bases_per_seq, k);
    int t, rc = ref_query_check(n_seqs, bases_per_seq, k, prefix_len, pattern);
    if (rc != REF_OK) return rc;
    if (threads < 1) threads = 1;
    if (threads == 1 || (n_seqs == 1 && rows < (uint64_t)threads * 1024))
        return ref_count_range(words, 0, n_seqs, bases_per_seq, stride_words, 0, UINT64_MAX, k,
                               prefix_bits, prefix_len, pattern, faithful, agg);
    jobs = (ref_mt_job *)calloc((size_t)threads, sizeof(*jobs));
    tids = (pthread_t *)calloc((size_t)threads, sizeof(*tids));
    for (t = 0; t < threads; t++) {
        ref_mt_job *j = &jobs[t];
        j->words = words;
        j->bases_per_seq = bases_per_seq;
        j->stride_words = stride_words;
        j->k = k;
        j->prefix_bits = prefix_bits;
        j->prefix_len = prefix_len;
        j->pattern = pattern;
        j->faithful = faithful;
        if (n_seqs == 1) { /* split one long sequence by k-mer rows */
            j->seq_first = 0;
            j->seq_last = 1;
            j->row_first = rows * (uint64_t)t / (uint64_t)threads;
            j->row_last = rows * (uint64_t)(t + 1) / (uint64_t)threads;
        } else { /* split a batch by sequences */
            j->seq_first = n_seqs * (uint64_t)t / (uint64_t)threads;
            j->seq_last = n_seqs * (uint64_t)(t + 1) / (uint64_t)threads;
            j->row_first = 0;
            j->row_last = UINT64_MAX;
        }
        j->agg = ref_agg_new((n_seqs * rows) / (uint64_t)threads + 1024);
        pthread_create(&tids[t], NULL, ref_mt_main, j);
    }
    agg->k = k;
    for (t = 0; t < threads; t++) {
        pthread_join(tids[t], NULL);
        if (jobs[t].rc != REF_OK) rc = jobs[t].rc;
    }
    /* phase 2: merge in parallel into `threads` disjoint shards of the caller's aggregate */
    if (agg->n_shards == 0) {
        agg->shards = (ref_agg **)calloc((size_t)threads, sizeof(ref_agg *));
        agg->n_shards = threads;
        for (t = 0; t < threads; t++) {
            agg->shards[t] = ref_agg_new((n_seqs * rows) / (uint64_t)threads + 1024);
            agg->shards[t]->k = k;
        }
    }
    if (agg->n_shards == threads) {
        for (t = 0; t < threads; t++) {
            jobs[t].all = jobs;
            jobs[t].n_jobs = threads;
            jobs[t].shard = t;
            jobs[t].final = agg->shards[t];
            pthread_create(&tids[t], NULL, ref_mt_merge, &jobs[t]);
        }
        for (t = 0; t < threads; t++) pthread_join(tids[t], NULL);
    } else { /* shard count differs from an earlier call: plain serial merge */
        for (t = 0; t < threads; t++) {
            uint64_t i;
            for (i = 0; i < jobs[t].agg->cap; i++)
                if (jobs[t].agg->slots[i].count)
                    ref_agg_add(agg, jobs[t].agg->slots[i].key, jobs[t].agg->slots[i].count);
        }
    }
    for (t = 0; t < threads; t++) ref_agg_free(jobs[t].agg);
    free(jobs);
    free(tids);
    return rc;
}

/* ---- the same query for inputs whose grouped result does not fit in memory ----------------------- */
/* SURVEY.md section 7, hard part 4: the 3.1 Gbp / k = 31 result is ~ 50 GB as a sorted list.  The rows are
 * produced exactly as in ref_count_range (window form or the faithful per-k-mer decode, same WHERE
 * evaluation), but grouped in `passes` x P disjoint hash partitions: pass g keeps the rows whose
 * partition hash falls in group g, lays them out by sub-partition, and every sub-partition is then
 * aggregated by the oracle's own hash aggregate (ref_agg: ref_kmer_hash + ref_kmer_eq).  Partitions
 * hold disjoint key sets, so total / distinct / unique add and the order-independent digest combines
 * (+ for the sums, ^ for the xors).  Memory: rows / passes * 8 bytes. */
typedef struct ref_big_job {
    const uint64_t *words;
    uint64_t n_words_seq, rows_per_seq, stride_words, row_first, row_last; /* global rows [first, last) */
    int k, prefix_len, faithful, pass, passes, rc;
    uint64_t prefix_bits, sub_mask;
    const char *pattern;
    uint64_t *hist;     /* P counters of this thread */
    uint64_t *cursor;   /* P write positions of this thread (scatter phase) */
    uint64_t *keys;     /* the pass's key array */
    /* aggregate phase */
    const uint64_t *sub_off;
    uint64_t n_sub;
    volatile uint64_t *next_sub;
    uint64_t stats[3], digest[4];
} ref_big_job;

static inline uint64_t ref_big_hash(uint64_t key) { return dnagpu_splitmix64(key ^ 0x6A09E667F3BCC909ull); }

/* walk the rows of one thread; mode 0 = histogram, 1 = scatter */
static int ref_big_walk(ref_big_job *j, int mode)
{
    uint64_t r = j->row_first;
    uint64_t s = j->rows_per_seq ? r / j->rows_per_seq : 0, i = j->rows_per_seq ? r % j->rows_per_seq : 0;
    int err;
    for (; r < j->row_last; r++) {
        const uint64_t *w = j->words + s * j->stride_words;
        uint64_t bits, h;
        if (j->faithful) {
            int rc = ref_generate_one(w, i, j->k, &bits);
            if (rc != REF_OK) return rc;
        } else {
            bits = ref_window(w, j->n_words_seq, i, j->k);
        }
        if (++i == j->rows_per_seq) {
            i = 0;
            s++;
        }
        if (j->prefix_len > 0 && !ref_starts_with(bits, j->k, j->prefix_bits, j->prefix_len, &err)) continue;
        if (j->pattern != NULL && !ref_contains(j->pattern, bits, j->k, &err)) continue;
        h = ref_big_hash(bits);
        if ((int)((h >> 40) % (uint64_t)j->passes) != j->pass) continue;
        if (mode == 0)
            j->hist[h & j->sub_mask]++;
        else
            j->keys[j->cursor[h & j->sub_mask]++] = bits;
    }
    return REF_OK;
}

static void *ref_big_hist_main(void *arg)
{
    ref_big_job *j = (ref_big_job *)arg;
    j->rc = ref_big_walk(j, 0);
    return NULL;
}
static void *ref_big_scatter_main(void *arg)
{
    ref_big_job *j = (ref_big_job *)arg;
    j->rc = ref_big_walk(j, 1);
    return NULL;
}
static void *ref_big_agg_main(void *arg)
{
    ref_big_job *j = (ref_big_job *)arg;
    for (;;) {
        uint64_t sub = __sync_fetch_and_add(j->next_sub, 1), beg, end, q, t, d, u, dg[4];
        ref_agg *a;
        if (sub >= j->n_sub) break;
        beg = j->sub_off[sub];
        end = j->sub_off[sub + 1];
        if (end == beg) continue;
        a = ref_agg_new((end - beg) < 1024 ? 1024 : (end - beg));
        if (!a) {
            j->rc = -1;
            break;
        }
        a->k = j->k;
        for (q = beg; q < end; q++) ref_agg_add(a, j->keys[q], 1);
        ref_agg_stats(a, &t, &d, &u);
        ref_agg_digest(a, dg);
        ref_agg_free(a);
        j->stats[0] += t;
        j->stats[1] += d;
        j->stats[2] += u;
        j->digest[0] += dg[0];
        j->digest[1] ^= dg[1];
        j->digest[2] += dg[2];
        j->digest[3] ^= dg[3];
    }
    return NULL;
}

int ref_count_query_big(const uint64_t *words, uint64_t n_seqs, uint64_t bases_per_seq,
                        uint64_t stride_words, int k, uint64_t prefix_bits, int prefix_len,
                        const char *pattern, int faithful, int passes, int threads,
                        uint64_t stats[3], uint64_t digest[4])
{
    const uint64_t rows_per_seq = ref_kmer_rows(bases_per_seq, k), rows = n_seqs * rows_per_seq;
    uint64_t P = 1, sub, *hist, *cursor, *sub_off, next_sub;
    ref_big_job *jobs;
    pthread_t *tids;
    int t, g, rc = ref_query_check(n_seqs, bases_per_seq, k, prefix_len, pattern);
    stats[0] = stats[1] = stats[2] = 0;
    digest[0] = digest[1] = digest[2] = digest[3] = 0;
    if (rc != REF_OK) return rc;
    if (rows == 0) return REF_OK;
    if (passes < 1) passes = 1;
    if (threads < 1) threads = 1;
    while (P < (1u << 16) && rows / (uint64_t)passes / P > 32768) P <<= 1; /* sub-partitions of ~32 K rows */
    jobs = (ref_big_job *)calloc((size_t)threads, sizeof(*jobs));
    tids = (pthread_t *)calloc((size_t)threads, sizeof(*tids));
    hist = (uint64_t *)calloc((size_t)threads * P, sizeof(uint64_t));
    cursor = (uint64_t *)calloc((size_t)threads * P, sizeof(uint64_t));
    sub_off = (uint64_t *)calloc(P + 1, sizeof(uint64_t));
    if (!jobs || !tids || !hist || !cursor || !sub_off) return -1;
    for (g = 0; g < passes && rc == REF_OK; g++) {
        uint64_t n_pass = 0, *keys;
        memset(hist, 0, (size_t)threads * P * sizeof(uint64_t));
        for (t = 0; t < threads; t++) {
            ref_big_job *j = &jobs[t];
            memset(j, 0, sizeof(*j));
            j->words = words;
            j->n_words_seq = ref_dna_words(bases_per_seq);
            j->rows_per_seq = rows_per_seq;
            j->stride_words = stride_words;
            j->row_first = rows / (uint64_t)threads * (uint64_t)t;
            j->row_last = t == threads - 1 ? rows : rows / (uint64_t)threads * (uint64_t)(t + 1);
            j->k = k;
            j->prefix_bits = prefix_bits;
            j->prefix_len = prefix_len;
            j->pattern = pattern;
            j->faithful = faithful;
            j->pass = g;
            j->passes = passes;
            j->sub_mask = P - 1;
            j->hist = hist + (size_t)t * P;
            j->cursor = cursor + (size_t)t * P;
            pthread_create(&tids[t], NULL, ref_big_hist_main, j);
        }
        for (t = 0; t < threads; t++) {
            pthread_join(tids[t], NULL);
            if (jobs[t].rc != REF_OK) rc = jobs[t].rc;
        }
        if (rc != REF_OK) break;
        for (sub = 0; sub < P; sub++) { /* sub-partition major, thread minor: deterministic layout */
            sub_off[sub] = n_pass;
            for (t = 0; t < threads; t++) {
                cursor[(size_t)t * P + sub] = n_pass;
                n_pass += hist[(size_t)t * P + sub];
            }
        }
        sub_off[P] = n_pass;
        keys = (uint64_t *)malloc((n_pass ? n_pass : 1) * sizeof(uint64_t));
        if (!keys) {
            rc = -1;
            break;
        }
        for (t = 0; t < threads; t++) {
            jobs[t].keys = keys;
            pthread_create(&tids[t], NULL, ref_big_scatter_main, &jobs[t]);
        }
        for (t = 0; t < threads; t++) pthread_join(tids[t], NULL);
        next_sub = 0;
        for (t = 0; t < threads; t++) {
            jobs[t].sub_off = sub_off;
            jobs[t].n_sub = P;
            jobs[t].next_sub = &next_sub;
            pthread_create(&tids[t], NULL, ref_big_agg_main, &jobs[t]);
        }
        for (t = 0; t < threads; t++) {
            pthread_join(tids[t], NULL);
            if (jobs[t].rc != REF_OK) rc = jobs[t].rc;
            stats[0] += jobs[t].stats[0];
            stats[1] += jobs[t].stats[1];
            stats[2] += jobs[t].stats[2];
            digest[0] += jobs[t].digest[0];
            digest[1] ^= jobs[t].digest[1];
            digest[2] += jobs[t].digest[2];
            digest[3] ^= jobs[t].digest[3];
        }
        free(keys);
    }
    free(jobs);
    free(tids);
    free(hist);
    free(cursor);
    free(sub_off);
    return rc;
}

/* ---- synthetic inputs ---------------------------------------------------------------------------- */
void ref_synth_seq(uint64_t seed, uint32_t repeat_every, uint64_t n_bases, uint64_t first_word,
                   uint64_t n_words, uint64_t *words)
{
    uint64_t j;
    for (j = 0; j < n_words; j++)
        words[j] = dnagpu_synth_seq_word(seed, repeat_every, n_bases, first_word + j);
}

void ref_synth_reads(uint64_t seed, uint32_t repeat_every, uint64_t first_read, uint64_t n_reads,
                     uint32_t bases_per_read, uint32_t stride_words, uint64_t *words)
{
    uint64_t r;
    uint32_t t;
    for (r = 0; r < n_reads; r++)
        for (t = 0; t < stride_words; t++)
            words[r * stride_words + t] = dnagpu_synth_read_word(seed, repeat_every, bases_per_read,
                                                                 stride_words, first_read + r, t);
}
