/*
 * dna_host.c -- see dna_host.h.  Host C over the libdnagpu C ABI.
 */
#include "dna_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct dnah_session {
    dnagpu_ctx *ctx;
};

static __thread char g_error[256];
const char *dnah_last_error(void) { return g_error; }
static int fail(const char *msg)
{
    snprintf(g_error, sizeof g_error, "%s", msg);
    return 1;
}
static int fail_gpu(dnagpu_ctx *ctx, int rc)
{
    const char *m = ctx ? dnagpu_last_error(ctx) : dnagpu_last_error(NULL);
    snprintf(g_error, sizeof g_error, "%s", (m && *m) ? m : dnagpu_strerror(rc));
    return rc;
}

/* ---- scalar glue ---- */
Dna *dna_make(const char *sequence)
{
    size_t n, i, words;
    Dna *dna;
    if (sequence == NULL || *sequence == '\0') { /* dna.c:160-161 */
        fail("DNA sequence cannot be empty");
        return NULL;
    }
    n = strlen(sequence);
    for (i = 0; i < n; i++) {
        char c = sequence[i];
        if (c != 'A' && c != 'T' && c != 'C' && c != 'G') { /* dna.c:165-166 */
            snprintf(g_error, sizeof g_error, "Invalid character in DNA sequence: %c", c);
            return NULL;
        }
    }
    words = (n * 2 + 63) / 64; /* dna.c:179-181 */
    dna = (Dna *)calloc(1, sizeof(Dna) + words * sizeof(uint64_t));
    if (!dna) {
        fail("out of memory");
        return NULL;
    }
    dna->length = n;
    for (i = 0; i < n; i++) { /* dna.c:114-128 */
        uint64_t code = sequence[i] == 'A' ? 0 : sequence[i] == 'T' ? 1 : sequence[i] == 'C' ? 2 : 3;
        dna->bit_sequence[i / 32] |= code << ((i * 2) % 64);
    }
    return dna;
}

char *dna_to_str(const Dna *dna)
{
    static const char letters[4] = {'A', 'T', 'C', 'G'};
    char *s = (char *)malloc(dna->length + 1);
    uint64_t i;
    if (!s) return NULL;
    for (i = 0; i < dna->length; i++) s[i] = letters[(dna->bit_sequence[i / 32] >> ((i * 2) % 64)) & 3];
    s[dna->length] = '\0';
    return s;
}

void dna_free(Dna *dna) { free(dna); }

int kmer_make(const char *sequence, Kmer *out)
{
    size_t n, i;
    uint64_t bits = 0;
    if (sequence == NULL) return fail("K-mer sequence cannot be NULL");          /* dna.c:495-496 */
    if (*sequence == '\0') return fail("K-mer sequence cannot be empty");        /* dna.c:460-461 */
    n = strlen(sequence);
    if (n > 32) return fail("K-mer length cannot exceed 32 nucleotides");        /* dna.c:466-467 */
    for (i = 0; i < n; i++) {
        uint64_t code;
        switch (sequence[i]) {
        case 'A': case 'X': code = 0; break; /* 'X' is the SP-GiST dummy, dna.c:413 */
        case 'T': code = 1; break;
        case 'C': code = 2; break;
        case 'G': code = 3; break;
        default:
            snprintf(g_error, sizeof g_error, "Invalid character in K-mer sequence: '%c'", sequence[i]);
            return 1;
        }
        bits |= code << (2 * i);
    }
    out->length = (int32_t)n;
    out->bit_sequence = bits;
    return 0;
}

int kmer_to_str(const Kmer *kmer, char out[33])
{
    static const char letters[4] = {'A', 'T', 'C', 'G'};
    int i;
    if (kmer->length <= 0 || kmer->length > 32)
        return fail("K-mer length must be between 1 and 32 nucleotides"); /* dna.c:433-435 */
    for (i = 0; i < kmer->length; i++) out[i] = letters[(kmer->bit_sequence >> (2 * i)) & 3];
    out[kmer->length] = '\0';
    return 0;
}

int qkmer_make(const char *sequence, Qkmer *out)
{
    size_t n, i;
    if (sequence == NULL || *sequence == '\0') return fail("qkmer pattern cannot be empty"); /* dna.c:877-879 */
    n = strlen(sequence);
    if (n > 32) return fail("Qkmer pattern length cannot exceed 32 characters");           /* dna.c:883-885 */
    for (i = 0; i < n; i++)
        if (!strchr("ATCGUWSMKRYBDHVN", sequence[i])) {                                    /* dna.c:888-895 */
            snprintf(g_error, sizeof g_error, "Invalid character in qkmer pattern: %c", sequence[i]);
            return 1;
        }
    memcpy(out->sequence, sequence, n + 1);
    return 0;
}

int kmer_eq_internal(const Kmer *a, const Kmer *b)
{
    return a->length == b->length && a->bit_sequence == b->bit_sequence;
}

/* ---- GPU-backed path ---- */
int dnah_open(dnah_session **out, int device)
{
    dnah_session *s = (dnah_session *)calloc(1, sizeof(*s));
    int rc;
    *out = NULL;
    if (!s) return fail("out of memory");
    rc = dnagpu_create(&s->ctx, device);
    if (rc != DNAGPU_OK) {
        free(s);
        return fail_gpu(NULL, rc);
    }
    *out = s;
    return 0;
}

void dnah_close(dnah_session *s)
{
    if (!s) return;
    dnagpu_destroy(s->ctx);
    free(s);
}

static void fill_where(dnagpu_where *w, const Kmer *prefix, const Qkmer *pattern)
{
    memset(w, 0, sizeof *w);
    if (prefix) {
        w->prefix_bits = prefix->bit_sequence;
        w->prefix_len = prefix->length;
    }
    if (pattern) w->qkmer = pattern->sequence;
}

static int rows_from_bits(const uint64_t *bits, uint64_t n, int k, Kmer **rows)
{
    uint64_t i;
    Kmer *r = (Kmer *)malloc((n ? n : 1) * sizeof(Kmer));
    if (!r) return fail("out of memory");
    for (i = 0; i < n; i++) {
        r[i].length = k;
        r[i].bit_sequence = bits[i];
    }
    *rows = r;
    return 0;
}

int generate_kmers_where(dnah_session *s, const Dna *dna, int k, const Kmer *prefix, const Qkmer *pattern,
                         Kmer **rows, uint64_t *n_rows)
{
    dnagpu_where w;
    uint64_t need = 0, *bits = NULL;
    int rc;
    *rows = NULL;
    *n_rows = 0;
    fill_where(&w, prefix, pattern);
    rc = dnagpu_filter_kmers(s->ctx, dna->bit_sequence, dna->length, k, &w, NULL, 0, &need);
    if (rc != DNAGPU_OK && rc != DNAGPU_ECAPACITY) return fail_gpu(s->ctx, rc);
    bits = (uint64_t *)malloc((need ? need : 1) * sizeof(uint64_t));
    if (!bits) return fail("out of memory");
    if (need) {
        rc = dnagpu_filter_kmers(s->ctx, dna->bit_sequence, dna->length, k, &w, bits, need, &need);
        if (rc != DNAGPU_OK) {
            free(bits);
            return fail_gpu(s->ctx, rc);
        }
    }
    rc = rows_from_bits(bits, need, k, rows);
    free(bits);
    *n_rows = need;
    return rc;
}

int generate_kmers(dnah_session *s, const Dna *dna, int k, Kmer **rows, uint64_t *n_rows)
{
    return generate_kmers_where(s, dna, k, NULL, NULL, rows, n_rows);
}

int count_kmers(dnah_session *s, const Dna *dna, int k, const Kmer *prefix, const Qkmer *pattern,
                KmerCount **rows, uint64_t *n_rows, dnagpu_stats *stats)
{
    dnagpu_where w;
    dnagpu_table *table = NULL;
    dnagpu_stats st;
    int rc;
    fill_where(&w, prefix, pattern);
    rc = dnagpu_count_kmers(s->ctx, dna->bit_sequence, dna->length, k, &w, &st, rows ? &table : NULL);
    if (rc != DNAGPU_OK) return fail_gpu(s->ctx, rc);
    if (stats) *stats = st;
    if (rows) {
        uint64_t n = dnagpu_table_rows(table), i;
        uint64_t *kk = (uint64_t *)malloc((n ? n : 1) * 8), *cc = (uint64_t *)malloc((n ? n : 1) * 8);
        KmerCount *r = (KmerCount *)malloc((n ? n : 1) * sizeof(KmerCount));
        if (!kk || !cc || !r) rc = fail("out of memory");
        if (rc == 0) rc = dnagpu_table_fetch(s->ctx, table, 0, n, kk, cc);
        if (rc == 0) {
            for (i = 0; i < n; i++) {
                r[i].kmer.length = k;
                r[i].kmer.bit_sequence = kk[i];
                r[i].count = cc[i];
            }
            *rows = r;
            *n_rows = n;
        } else {
            if (rc != 1) fail_gpu(s->ctx, rc);
            free(r);
        }
        free(kk);
        free(cc);
        dnagpu_table_free(table);
    }
    return rc;
}
