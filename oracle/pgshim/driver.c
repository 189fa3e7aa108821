/*
 * pgshim/driver.c -- calls the reference's own fmgr functions (dna.c, compiled unmodified)
 * the way the PostgreSQL executor does, and exposes the results through a plain C ABI for
 * the tests and for bench.py's reference arm.  TEST INFRASTRUCTURE ONLY.
 *
 *   FunctionScan   -> the ValuePerCall loop over generate_kmers (dna.c:743-837)
 *   qual           -> starts_with (dna.c:842-866) / contains (dna.c:1091-1135) per row
 *   HashAggregate  -> a table whose bucket comes from kmer_hash (dna.c:722-735) and whose
 *                     matches are confirmed by kmer_eq (dna.c:686-696), count(*) per group
 *
 * The value layouts are the reference's structs (dna.c:42-47, 61-65, 81-84), restated here
 * because dna.c keeps them private.
 */
#include "postgres.h"

#include <pthread.h>

typedef struct RefDna { /* dna.c:42-47 */
    char vl_len_[4];
    uint64_t length;
    uint64_t bit_sequence[];
} RefDna;
typedef struct RefKmer { /* dna.c:61-65 */
    int32 length;
    uint64_t bit_sequence;
} RefKmer;
typedef struct RefQkmer { /* dna.c:81-84 */
    char vl_len_[4];
    char sequence[];
} RefQkmer;

extern Datum dna_in(PG_FUNCTION_ARGS);
extern Datum dna_out(PG_FUNCTION_ARGS);
extern Datum kmer_in(PG_FUNCTION_ARGS);
extern Datum kmer_out(PG_FUNCTION_ARGS);
extern Datum qkmer_in(PG_FUNCTION_ARGS);
extern Datum generate_kmers(PG_FUNCTION_ARGS);
extern Datum starts_with(PG_FUNCTION_ARGS);
extern Datum contains(PG_FUNCTION_ARGS);
extern Datum kmer_hash(PG_FUNCTION_ARGS);
extern Datum kmer_eq(PG_FUNCTION_ARGS);

extern __thread char shim_error_text[512];
extern __thread jmp_buf *shim_error_jmp;

static Datum call(Datum (*fn)(PG_FUNCTION_ARGS), FmgrInfo *fl, ReturnSetInfo *rsi, bool *isnull, int nargs,
                  Datum a0, Datum a1)
{
    FunctionCallInfoBaseData fc;
    Datum r;
    memset(&fc, 0, sizeof fc);
    fc.flinfo = fl;
    fc.resultinfo = rsi;
    fc.nargs = (short)nargs;
    fc.args[0].value = a0;
    fc.args[1].value = a1;
    r = fn(&fc);
    if (isnull) *isnull = fc.isnull;
    return r;
}
#define CALL1(fn, a) call(fn, &fl_, NULL, NULL, 1, (Datum)(a), 0)
#define CALL2(fn, a, b) call(fn, &fl_, NULL, NULL, 2, (Datum)(a), (Datum)(b))

/* run `body` with ereport(ERROR) turned into `return 1` + message */
#define GUARDED(err, errcap, ...)                                         \
    do {                                                                  \
        jmp_buf jb_;                                                      \
        FmgrInfo fl_;                                                     \
        memset(&fl_, 0, sizeof fl_);                                      \
        (void)fl_;                                                        \
        shim_error_jmp = &jb_;                                            \
        if (setjmp(jb_) != 0) {                                           \
            shim_error_jmp = NULL;                                        \
            shim_abort_cleanup(); /* AbortTransaction: contexts reset */  \
            if (err) snprintf(err, errcap, "%s", shim_error_text);        \
            return 1;                                                     \
        }                                                                 \
        __VA_ARGS__;                                                      \
        shim_error_jmp = NULL;                                            \
    } while (0)

static RefDna *dna_from_words(const uint64_t *words, uint64_t n_bases)
{
    uint64_t nw = (n_bases * 2 + 63) / 64; /* dna.c:179-181 */
    Size sz = offsetof(RefDna, bit_sequence) + nw * sizeof(uint64_t);
    RefDna *d = (RefDna *)palloc0(sz);
    SET_VARSIZE(d, sz);
    d->length = n_bases;
    memcpy(d->bit_sequence, words, nw * sizeof(uint64_t));
    return d;
}

/* ---- scalar I/O ---- */
int dnaref_dna_in(const char *text, uint64_t *words, uint64_t cap_words, uint64_t *n_bases, char *err, size_t errcap)
{
    GUARDED(err, errcap, {
        RefDna *d = (RefDna *)DatumGetPointer(CALL1(dna_in, text));
        uint64_t nw = (d->length * 2 + 63) / 64;
        *n_bases = d->length;
        if (nw > cap_words) {
            snprintf(shim_error_text, sizeof shim_error_text, "driver: word buffer too small");
            longjmp(jb_, 1);
        }
        memcpy(words, d->bit_sequence, nw * 8);
        pfree(d);
    });
    return 0;
}

int dnaref_dna_out(const uint64_t *words, uint64_t n_bases, char *out, size_t cap, char *err, size_t errcap)
{
    GUARDED(err, errcap, {
        RefDna *d = dna_from_words(words, n_bases);
        char *s = DatumGetCString(CALL1(dna_out, d));
        snprintf(out, cap, "%s", s);
        pfree(s);
        pfree(d);
    });
    return 0;
}

int dnaref_kmer_in(const char *text, uint64_t *bits, int32_t *length, char *err, size_t errcap)
{
    GUARDED(err, errcap, {
        RefKmer *k = (RefKmer *)DatumGetPointer(CALL1(kmer_in, text));
        *bits = k->bit_sequence;
        *length = k->length;
        pfree(k);
    });
    return 0;
}

int dnaref_kmer_out(uint64_t bits, int32_t length, char *out, size_t cap, char *err, size_t errcap)
{
    GUARDED(err, errcap, {
        RefKmer k;
        char *s;
        k.length = length;
        k.bit_sequence = bits;
        s = DatumGetCString(CALL1(kmer_out, &k));
        snprintf(out, cap, "%s", s);
        pfree(s);
    });
    return 0;
}

int dnaref_qkmer_in(const char *text, char *err, size_t errcap)
{
    GUARDED(err, errcap, {
        void *q = DatumGetPointer(CALL1(qkmer_in, text));
        pfree(q);
    });
    return 0;
}

/* ---- operators ---- */
int dnaref_starts_with(uint64_t kbits, int32_t klen, uint64_t pbits, int32_t plen, int *result, char *err,
                       size_t errcap)
{
    GUARDED(err, errcap, {
        RefKmer k, p;
        k.length = klen;
        k.bit_sequence = kbits;
        p.length = plen;
        p.bit_sequence = pbits;
        *result = DatumGetBool(CALL2(starts_with, &k, &p));
    });
    return 0;
}

static RefQkmer *qkmer_from_text(const char *pattern)
{ /* what qkmer_make builds (dna.c:917-927), without its validation: contains() trusts its input */
    size_t n = strlen(pattern);
    RefQkmer *q = (RefQkmer *)palloc0(offsetof(RefQkmer, sequence) + n + 1);
    SET_VARSIZE(q, offsetof(RefQkmer, sequence) + n + 1);
    memcpy(q->sequence, pattern, n + 1);
    return q;
}

int dnaref_contains(const char *pattern, uint64_t kbits, int32_t klen, int *result, char *err, size_t errcap)
{
    GUARDED(err, errcap, {
        RefKmer k;
        RefQkmer *q = (RefQkmer *)DatumGetPointer(CALL1(qkmer_in, pattern)); /* validates like the server */
        k.length = klen;
        k.bit_sequence = kbits;
        *result = DatumGetBool(CALL2(contains, q, &k));
        pfree(q);
    });
    return 0;
}

uint32_t dnaref_kmer_hash(uint64_t bits)
{
    RefKmer k;
    FmgrInfo fl_;
    memset(&fl_, 0, sizeof fl_);
    k.length = 0;
    k.bit_sequence = bits;
    return DatumGetUInt32(CALL1(kmer_hash, &k));
}

int dnaref_kmer_eq(uint64_t abits, int32_t alen, uint64_t bbits, int32_t blen)
{
    RefKmer a, b;
    FmgrInfo fl_;
    memset(&fl_, 0, sizeof fl_);
    a.length = alen;
    a.bit_sequence = abits;
    b.length = blen;
    b.bit_sequence = bbits;
    return DatumGetBool(CALL2(kmer_eq, &a, &b));
}

/* ---- FunctionScan over generate_kmers, with optional quals ---- */
typedef int (*row_fn)(void *arg, const RefKmer *k);

/* The reference is undefined for length < k - 1 (unsigned wrap, dna.c:781) and for
 * >= 2^30 bases (int overflow, dna.c:805-807): the driver does not call it there. */
static int scan_generate_kmers(const RefDna *d, int k, const RefKmer *prefix, const RefQkmer *q, row_fn fn,
                               void *arg)
{
    FmgrInfo fl;
    ReturnSetInfo rsi;
    memset(&fl, 0, sizeof fl);
    if (k >= 1 && k <= 32 && d->length + 1 < (uint64_t)k) return 0;
    for (;;) {
        bool isnull = false;
        Datum r;
        rsi.isDone = ExprSingleResult;
        r = call(generate_kmers, &fl, &rsi, &isnull, 2, PointerGetDatum(d), Int32GetDatum(k));
        if (rsi.isDone == ExprEndResult) break;
        {
            RefKmer *km = (RefKmer *)DatumGetPointer(r);
            FmgrInfo fl2;
            int keep = 1;
            memset(&fl2, 0, sizeof fl2);
            if (prefix && !DatumGetBool(call(starts_with, &fl2, NULL, NULL, 2, PointerGetDatum(km), PointerGetDatum(prefix))))
                keep = 0;
            if (keep && q && !DatumGetBool(call(contains, &fl2, NULL, NULL, 2, PointerGetDatum(q), PointerGetDatum(km))))
                keep = 0;
            if (keep && fn(arg, km) != 0) {
                pfree(km);
                return 1;
            }
            pfree(km); /* the executor's per-tuple context reset */
        }
    }
    return 0;
}

typedef struct collect {
    uint64_t *out;
    uint64_t cap, n;
    int k, bad_len;
} collect;
static int collect_row(void *arg, const RefKmer *km)
{
    collect *c = (collect *)arg;
    if (km->length != c->k) c->bad_len = 1;
    if (c->n < c->cap) c->out[c->n] = km->bit_sequence;
    c->n++;
    return 0;
}

int dnaref_generate_kmers(const uint64_t *words, uint64_t n_bases, int k, uint64_t prefix_bits, int32_t prefix_len,
                          const char *pattern, uint64_t *out, uint64_t cap, uint64_t *n_out, char *err,
                          size_t errcap)
{
    if (n_bases >= (1ull << 30)) {
        if (err) snprintf(err, errcap, "driver: the reference's int indices overflow at 2^30 bases");
        return 2;
    }
    GUARDED(err, errcap, {
        RefDna *d = dna_from_words(words, n_bases);
        RefKmer pk;
        RefQkmer *q = pattern ? (RefQkmer *)DatumGetPointer(CALL1(qkmer_in, pattern)) : NULL;
        collect c;
        c.out = out;
        c.cap = cap;
        c.n = 0;
        c.k = k;
        c.bad_len = 0;
        pk.length = prefix_len;
        pk.bit_sequence = prefix_bits;
        scan_generate_kmers(d, k, prefix_len > 0 ? &pk : NULL, q, collect_row, &c);
        *n_out = c.n;
        if (q) pfree(q);
        pfree(d);
        if (c.bad_len) {
            snprintf(shim_error_text, sizeof shim_error_text, "driver: a row's Kmer.length differs from k");
            longjmp(jb_, 1);
        }
    });
    return 0;
}

/* ---- HashAggregate: GROUP BY kmer, count(*) ---- */
typedef struct agg_slot {
    uint64_t key, count; /* count 0 = empty */
} agg_slot;
typedef struct agg_table {
    agg_slot *slots;
    uint64_t cap, groups;
    int k;
} agg_table;

static void agg_init(agg_table *t, uint64_t expected, int k)
{
    uint64_t cap = 1024;
    while (cap < expected * 2) cap <<= 1;
    t->slots = (agg_slot *)calloc(cap, sizeof(agg_slot));
    t->cap = cap;
    t->groups = 0;
    t->k = k;
}
static void agg_add(agg_table *t, uint64_t key, uint64_t times);
static void agg_grow(agg_table *t)
{
    agg_table n;
    uint64_t i;
    agg_init(&n, t->cap, t->k);
    for (i = 0; i < t->cap; i++)
        if (t->slots[i].count) agg_add(&n, t->slots[i].key, t->slots[i].count);
    free(t->slots);
    *t = n;
}
static void agg_add(agg_table *t, uint64_t key, uint64_t times)
{
    uint64_t i;
    if ((t->groups + 1) * 4 > t->cap * 3) agg_grow(t);
    i = dnaref_kmer_hash(key) & (t->cap - 1);                /* the opclass hash function */
    for (;;) {
        agg_slot *s = &t->slots[i];
        if (s->count == 0) {
            s->key = key;
            s->count = times;
            t->groups++;
            return;
        }
        if (dnaref_kmer_eq(s->key, t->k, key, t->k)) {        /* the opclass equality operator */
            s->count += times;
            return;
        }
        i = (i + 1) & (t->cap - 1);
    }
}
static int agg_row(void *arg, const RefKmer *km)
{
    agg_add((agg_table *)arg, km->bit_sequence, 1);
    return 0;
}

typedef struct job {
    const uint64_t *words;
    uint64_t seq_first, seq_last, bases_per_seq, stride;
    int k, failed, n_jobs, shard;
    uint64_t prefix_bits;
    int32_t prefix_len;
    const char *pattern;
    agg_table table, merged;
    struct job *all;
    char err[256];
} job;

static int job_scan(job *j)
{
    GUARDED(j->err, sizeof j->err, {
        RefQkmer *q = j->pattern ? (RefQkmer *)DatumGetPointer(CALL1(qkmer_in, j->pattern)) : NULL;
        RefKmer pk;
        uint64_t s;
        pk.length = j->prefix_len;
        pk.bit_sequence = j->prefix_bits;
        for (s = j->seq_first; s < j->seq_last; s++) {
            RefDna *d = dna_from_words(j->words + s * j->stride, j->bases_per_seq);
            scan_generate_kmers(d, j->k, j->prefix_len > 0 ? &pk : NULL, q, agg_row, &j->table);
            pfree(d);
        }
        if (q) pfree(q);
    });
    return 0;
}
static void *job_main(void *arg)
{
    job *j = (job *)arg;
    j->failed = job_scan(j);
    return NULL;
}
static void *merge_main(void *arg)
{
    job *j = (job *)arg;
    int t;
    for (t = 0; t < j->n_jobs; t++) {
        const agg_table *a = &j->all[t].table;
        uint64_t i;
        for (i = 0; i < a->cap; i++)
            if (a->slots[i].count && (int)((dnaref_kmer_hash(a->slots[i].key) >> 8) % (uint32_t)j->n_jobs) == j->shard)
                agg_add(&j->merged, a->slots[i].key, a->slots[i].count);
    }
    return NULL;
}

/* mix used by the order-independent digest (same as ref_cpu.c / dnagpu_synth.h splitmix64) */
static uint64_t sm64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/*
 * SELECT count(*), sum(c), count(*) FILTER (WHERE c = 1) FROM (SELECT count(*) c FROM seqs,
 * generate_kmers(seq, k) AS g(kmer) [WHERE kmer ^@ prefix AND pattern @> kmer] GROUP BY kmer).
 * A batch of sequences is split over `threads` workers (PostgreSQL itself would run this plan
 * on one core: the functions are not PARALLEL SAFE); rows_kmers/rows_counts (optional, cap
 * entries) receive the groups sorted by kmer.
 */
static int pair_cmp(const void *a, const void *b)
{
    uint64_t x = ((const agg_slot *)a)->key, y = ((const agg_slot *)b)->key;
    return x < y ? -1 : (x > y ? 1 : 0);
}

int dnaref_count(const uint64_t *words, uint64_t n_seqs, uint64_t bases_per_seq, uint64_t stride, int k,
                 uint64_t prefix_bits, int32_t prefix_len, const char *pattern, int threads, uint64_t stats[3],
                 uint64_t digest[4], uint64_t *rows_kmers, uint64_t *rows_counts, uint64_t rows_cap,
                 uint64_t *n_rows, char *err, size_t errcap)
{
    job *jobs;
    pthread_t *tids;
    int t, rc = 0;
    uint64_t m = 0;
    agg_slot *rows = NULL;
    if (bases_per_seq >= (1ull << 30)) {
        if (err) snprintf(err, errcap, "driver: the reference's int indices overflow at 2^30 bases");
        return 2;
    }
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n_seqs) threads = (int)(n_seqs ? n_seqs : 1);
    jobs = (job *)calloc((size_t)threads, sizeof(job));
    tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (t = 0; t < threads; t++) {
        job *j = &jobs[t];
        uint64_t rows_est = bases_per_seq >= (uint64_t)k ? bases_per_seq - k + 1 : 0;
        j->words = words;
        j->seq_first = n_seqs * (uint64_t)t / (uint64_t)threads;
        j->seq_last = n_seqs * (uint64_t)(t + 1) / (uint64_t)threads;
        j->bases_per_seq = bases_per_seq;
        j->stride = stride;
        j->k = k;
        j->prefix_bits = prefix_bits;
        j->prefix_len = prefix_len;
        j->pattern = pattern;
        j->all = jobs;
        j->n_jobs = threads;
        j->shard = t;
        agg_init(&j->table, (j->seq_last - j->seq_first) * rows_est / ((prefix_len || pattern) ? 16 : 1) + 1024, k);
        pthread_create(&tids[t], NULL, job_main, j);
    }
    for (t = 0; t < threads; t++) {
        pthread_join(tids[t], NULL);
        if (jobs[t].failed && !rc) {
            rc = 1;
            if (err) snprintf(err, errcap, "%s", jobs[t].err);
        }
    }
    stats[0] = stats[1] = stats[2] = 0;
    digest[0] = digest[1] = digest[2] = digest[3] = 0;
    if (!rc) {
        for (t = 0; t < threads; t++) {
            agg_init(&jobs[t].merged, jobs[t].table.groups + 1024, k);
            pthread_create(&tids[t], NULL, merge_main, &jobs[t]);
        }
        for (t = 0; t < threads; t++) pthread_join(tids[t], NULL);
        for (t = 0; t < threads; t++) stats[1] += jobs[t].merged.groups;
        if (rows_kmers && rows_counts) rows = (agg_slot *)malloc((stats[1] ? stats[1] : 1) * sizeof(agg_slot));
        for (t = 0; t < threads; t++) {
            const agg_table *a = &jobs[t].merged;
            uint64_t i;
            for (i = 0; i < a->cap; i++) {
                uint64_t c = a->slots[i].count, key = a->slots[i].key, x, y;
                if (!c) continue;
                stats[0] += c;
                stats[2] += (c == 1);
                x = sm64(key) * c;
                y = sm64(key ^ (c * 0x9E3779B97F4A7C15ull));
                digest[0] += x;
                digest[1] ^= x;
                digest[2] += y;
                digest[3] ^= y;
                if (rows) rows[m++] = a->slots[i];
            }
        }
        if (rows) {
            uint64_t i;
            qsort(rows, m, sizeof(agg_slot), pair_cmp);
            for (i = 0; i < m && i < rows_cap; i++) {
                rows_kmers[i] = rows[i].key;
                rows_counts[i] = rows[i].count;
            }
            free(rows);
        }
        if (n_rows) *n_rows = stats[1];
    }
    for (t = 0; t < threads; t++) {
        free(jobs[t].table.slots);
        free(jobs[t].merged.slots);
    }
    free(jobs);
    free(tids);
    return rc;
}

#ifdef DNAREF_WITH_GLUE
/*
 * The GPU glue's functions (dna-sequences-pg-extension_b200/pg/dna_gpu.c), linked into this build together
 * with dna.c.  The driver plays the executor: it supplies the expected row type, passes SQL NULLs, runs the
 * SRF loop (optionally abandoning it half way, as a LIMIT or a cancelled query does), runs an aggregate's
 * transition / final functions over a table of values, and unpacks the composite Datums.
 */
extern Datum kmer_stats(PG_FUNCTION_ARGS);
extern Datum count_kmers(PG_FUNCTION_ARGS);
extern Datum generate_kmers_where(PG_FUNCTION_ARGS);
extern Datum kmer_stats_agg_trans(PG_FUNCTION_ARGS);
extern Datum kmer_stats_agg_final(PG_FUNCTION_ARGS);
extern int dna_gpu_live_tables(void);
extern int dna_gpu_device_count(void);

int dnaref_live_tables(void) { return dna_gpu_live_tables(); }
int dnaref_device_count(void) { return dna_gpu_device_count(); }
int dnaref_live_contexts(void) { return shim_live_contexts(); }

/* fn(dna, k) or fn(dna, k, prefix kmer | NULL, qkmer | NULL) */
static Datum call_glue(Datum (*fn)(PG_FUNCTION_ARGS), FmgrInfo *fl, ReturnSetInfo *rsi, TupleDesc desc, RefDna *d, int k,
                       int with_where, RefKmer *prefix, RefQkmer *q)
{
    FunctionCallInfoBaseData fc;
    memset(&fc, 0, sizeof fc);
    fc.flinfo = fl;
    fc.resultinfo = rsi;
    fc.nargs = with_where ? 4 : 2;
    fc.args[0].value = PointerGetDatum(d);
    fc.args[1].value = Int32GetDatum(k);
    if (with_where) {
        fc.args[2].value = PointerGetDatum(prefix);
        fc.args[2].isnull = prefix == NULL;
        fc.args[3].value = PointerGetDatum(q);
        fc.args[3].isnull = q == NULL;
    }
    fc.shim_result_desc = desc;
    return fn(&fc);
}

static RefKmer *prefix_arg(uint64_t prefix_bits, int32_t prefix_len, RefKmer *store)
{
    if (prefix_len <= 0) return NULL;
    store->length = prefix_len;
    store->bit_sequence = prefix_bits;
    return store;
}

/* SELECT * FROM kmer_stats(dna, k [, prefix, pattern]);  with_where = 0 calls the two-argument form */
int dnaref_kmer_stats(const uint64_t *words, uint64_t n_bases, int k, int with_where, uint64_t prefix_bits,
                      int32_t prefix_len, const char *pattern, int64_t stats[3], char *err, size_t errcap)
{
    GUARDED(err, errcap, {
        RefDna *d = dna_from_words(words, n_bases);
        RefKmer pk;
        RefQkmer *q = pattern ? qkmer_from_text(pattern) : NULL;
        TupleDescData desc = {3};
        HeapTuple t = (HeapTuple)DatumGetPointer(call_glue(kmer_stats, &fl_, NULL, &desc, d, k, with_where,
                                                           prefix_arg(prefix_bits, prefix_len, &pk), q));
        int i;
        for (i = 0; i < 3; i++) stats[i] = DatumGetInt64(t->values[i]);
        heap_freetuple(t);
        if (q) pfree(q);
        pfree(d);
    });
    return 0;
}

/* SELECT * FROM count_kmers(dna, k [, prefix, pattern]) [LIMIT stop_after]: rows in the order the function
 * returns them.  stop_after < the number of rows abandons the scan there, the way the executor does for a
 * LIMIT or a cancelled query: the SRF is never called again and its memory context is reset. */
int dnaref_count_kmers(const uint64_t *words, uint64_t n_bases, int k, int with_where, uint64_t prefix_bits,
                       int32_t prefix_len, const char *pattern, uint64_t stop_after, uint64_t *kmers, int64_t *counts,
                       uint64_t cap, uint64_t *n_out, char *err, size_t errcap)
{
    GUARDED(err, errcap, {
        RefDna *d = dna_from_words(words, n_bases);
        RefKmer pk;
        RefQkmer *q = pattern ? qkmer_from_text(pattern) : NULL;
        TupleDescData desc = {2};
        ReturnSetInfo rsi;
        uint64_t n = 0;
        int bad_len = 0;
        for (;;) {
            Datum r;
            if (n >= stop_after) { /* abandoned scan: ExecEndFunctionScan -> the SRF's context goes away */
                FuncCallContext *f = (FuncCallContext *)fl_.fn_extra;
                if (f) shim_end_MultiFuncCall(&(FunctionCallInfoBaseData){.flinfo = &fl_}, f);
                break;
            }
            rsi.isDone = ExprSingleResult;
            r = call_glue(count_kmers, &fl_, &rsi, &desc, d, k, with_where, prefix_arg(prefix_bits, prefix_len, &pk), q);
            if (rsi.isDone == ExprEndResult) break;
            {
                HeapTuple t = (HeapTuple)DatumGetPointer(r);
                RefKmer *km = (RefKmer *)DatumGetPointer(t->values[0]);
                if (km->length != k || t->natts != 2 || t->nulls[0] || t->nulls[1]) bad_len = 1;
                if (n < cap) {
                    kmers[n] = km->bit_sequence;
                    counts[n] = DatumGetInt64(t->values[1]);
                }
                n++;
                pfree(km);
                heap_freetuple(t);
            }
        }
        *n_out = n;
        if (q) pfree(q);
        pfree(d);
        if (bad_len) {
            snprintf(shim_error_text, sizeof shim_error_text, "driver: malformed count_kmers row");
            longjmp(jb_, 1);
        }
    });
    return 0;
}

/* SELECT * FROM generate_kmers_where(dna, k, prefix, pattern): rows in order */
int dnaref_generate_kmers_where(const uint64_t *words, uint64_t n_bases, int k, uint64_t prefix_bits, int32_t prefix_len,
                                const char *pattern, uint64_t *out, uint64_t cap, uint64_t *n_out, char *err,
                                size_t errcap)
{
    GUARDED(err, errcap, {
        RefDna *d = dna_from_words(words, n_bases);
        RefKmer pk;
        RefQkmer *q = pattern ? qkmer_from_text(pattern) : NULL;
        ReturnSetInfo rsi;
        uint64_t n = 0;
        for (;;) {
            Datum r;
            rsi.isDone = ExprSingleResult;
            r = call_glue(generate_kmers_where, &fl_, &rsi, NULL, d, k, 1, prefix_arg(prefix_bits, prefix_len, &pk), q);
            if (rsi.isDone == ExprEndResult) break;
            {
                RefKmer *km = (RefKmer *)DatumGetPointer(r);
                if (n < cap) out[n] = km->bit_sequence;
                n += km->length == k ? 1 : (1ull << 40); /* a wrong Kmer.length shows up as an absurd count */
                pfree(km);
            }
        }
        *n_out = n;
        if (q) pfree(q);
        pfree(d);
    });
    return 0;
}

/* SELECT (kmer_stats_agg(sequence, k)).* FROM t: values of n_bases[s] bases at words[word_off[s]]; a value with
 * n_bases[s] == UINT64_MAX is a SQL NULL */
int dnaref_kmer_stats_agg(const uint64_t *words, const uint64_t *word_off, const uint64_t *n_bases, uint64_t n_seqs, int k,
                          int64_t stats[3], char *err, size_t errcap)
{
    GUARDED(err, errcap, {
        ShimAggContext agg = {0x4147, shim_context_create()};
        TupleDescData desc = {3};
        Datum state = 0;
        bool state_null = true;
        uint64_t s;
        int i;
        HeapTuple t;
        for (s = 0; s < n_seqs; s++) {
            FunctionCallInfoBaseData fc;
            RefDna *d = n_bases[s] == UINT64_MAX ? NULL : dna_from_words(words + word_off[s], n_bases[s]);
            memset(&fc, 0, sizeof fc);
            fc.flinfo = &fl_;
            fc.context = &agg;
            fc.nargs = 3;
            fc.args[0].value = state;
            fc.args[0].isnull = state_null;
            fc.args[1].value = PointerGetDatum(d);
            fc.args[1].isnull = d == NULL;
            fc.args[2].value = Int32GetDatum(k);
            state = kmer_stats_agg_trans(&fc);
            state_null = fc.isnull;
            if (d) pfree(d);
        }
        {
            FunctionCallInfoBaseData fc;
            memset(&fc, 0, sizeof fc);
            fc.flinfo = &fl_;
            fc.context = &agg;
            fc.nargs = 1;
            fc.args[0].value = state;
            fc.args[0].isnull = state_null;
            fc.shim_result_desc = &desc;
            t = (HeapTuple)DatumGetPointer(kmer_stats_agg_final(&fc));
        }
        for (i = 0; i < 3; i++) stats[i] = DatumGetInt64(t->values[i]);
        heap_freetuple(t);
        shim_context_delete(agg.aggcontext);
    });
    return 0;
}
#endif /* DNAREF_WITH_GLUE */
