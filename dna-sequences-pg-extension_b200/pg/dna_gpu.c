/*
 * dna_gpu.c -- PostgreSQL-side glue: the fmgr entry points of the k-mer path re-pointed at
 * libdnagpu.  Built as part of the extension's .so in place of the corresponding functions
 * of the reference's dna.c (same symbol names, same SQL declarations, dna--1.0.sql:188-191),
 * so `CREATE EXTENSION dna` and every existing query keep working.
 *
 *   generate_kmers(dna, int) SETOF kmer   drop-in for dna.c:743-837.  The FIRST call extracts
 *       every k-mer on the GPU in one launch into multi_call_memory_ctx; the remaining calls
 *       hand the rows out one Datum at a time (the SRF protocol is row-at-a-time by design).
 *   kmer_stats(dna, int, OUT total, OUT distinct, OUT uniq)   new pushdown: the README's
 *       total / distinct / unique query (README.md:122-130) without ever materialising the
 *       k-mers in the executor.
 *   count_kmers(dna, int) SETOF (kmer, count)   new pushdown: the GROUP BY of README.md:107-116.
 *
 * There is no PostgreSQL in the build image (no pg_config, no server headers).  The tests
 * compile this file, together with the reference's dna.c, against the PostgreSQL API shim
 * they own and drive it through the fmgr / SRF protocol on the GPU (tests/test_gpu_pg_glue.py);
 * pg/Makefile builds the real extension on a machine that has PGXS.  INTEGRATION.md has the steps.
 */
#include "postgres.h"

#include "fmgr.h"
#include "funcapi.h"

#include "../../include/dnagpu.h"

#ifndef DNAGPU_GLUE_NO_MODULE_MAGIC /* dna.c already carries PG_MODULE_MAGIC when linked together */
PG_MODULE_MAGIC;
#endif

/* the reference's value layouts (dna.c:42-47, 61-65) */
typedef struct Dna {
    char vl_len_[4];
    uint64_t length;
    uint64_t bit_sequence[FLEXIBLE_ARRAY_MEMBER];
} Dna;
typedef struct Kmer {
    int32 length;
    uint64_t bit_sequence;
} Kmer;

/* One CUDA context per backend process, created on first use: contexts do not survive the
 * postmaster's fork, and most backends never touch a dna value. */
static dnagpu_ctx *backend_ctx = NULL;

static dnagpu_ctx *
gpu(void)
{
    if (backend_ctx == NULL)
    {
        int rc = dnagpu_create(&backend_ctx, 0);

        if (rc != DNAGPU_OK)
            ereport(ERROR, (errmsg("dnagpu: %s", dnagpu_last_error(NULL))));
    }
    return backend_ctx;
}

typedef struct GenerateState
{
    uint64_t   *bits;           /* all rows, extracted by the first call */
    int         k;
} GenerateState;

PG_FUNCTION_INFO_V1(generate_kmers);
Datum
generate_kmers(PG_FUNCTION_ARGS)
{
    FuncCallContext *funcctx;
    GenerateState *state;

    if (SRF_IS_FIRSTCALL())
    {
        MemoryContext oldcontext;
        Dna        *dna;
        int         k;
        uint64_t    rows, got = 0;
        int         rc;

        funcctx = SRF_FIRSTCALL_INIT();
        oldcontext = MemoryContextSwitchTo(funcctx->multi_call_memory_ctx);

        dna = (Dna *) PG_GETARG_VARLENA_P(0);
        k = PG_GETARG_INT32(1);
        if (k <= 0 || k > 32)   /* dna.c:772-773 */
            ereport(ERROR, (errmsg("Invalid k value: must be between 1 and 32")));

        /* dna.c:781 computes length - k + 1 unsigned; a sequence shorter than k has no rows */
        rows = dna->length >= (uint64_t) k ? dna->length - (uint64_t) k + 1 : 0;
        state = (GenerateState *) palloc(sizeof(GenerateState));
        state->k = k;
        state->bits = (uint64_t *) palloc(sizeof(uint64_t) * (rows ? rows : 1));
        /* the varlena may be 4-byte aligned (dna--1.0.sql:31): the library copies bytewise */
        rc = dnagpu_generate_kmers(gpu(), dna->bit_sequence, dna->length, k, state->bits, rows, &got);
        if (rc != DNAGPU_OK)
            ereport(ERROR, (errmsg("%s", dnagpu_last_error(backend_ctx))));

        funcctx->user_fctx = state;
        funcctx->max_calls = got;
        MemoryContextSwitchTo(oldcontext);
    }

    funcctx = SRF_PERCALL_SETUP();
    state = (GenerateState *) funcctx->user_fctx;

    if (funcctx->call_cntr < funcctx->max_calls)
    {
        Kmer       *kmer = (Kmer *) palloc0(sizeof(Kmer));

        kmer->length = state->k;
        kmer->bit_sequence = state->bits[funcctx->call_cntr];
        SRF_RETURN_NEXT(funcctx, PointerGetDatum(kmer));
    }
    SRF_RETURN_DONE(funcctx);
}

#include "access/htup_details.h"
#include "utils/builtins.h"

static void
check_k(int k)
{
    if (k <= 0 || k > 32)       /* dna.c:772-773, same text */
        ereport(ERROR, (errmsg("Invalid k value: must be between 1 and 32")));
}

PG_FUNCTION_INFO_V1(kmer_stats);
Datum
kmer_stats(PG_FUNCTION_ARGS)
{
    Dna        *dna = (Dna *) PG_GETARG_VARLENA_P(0);
    int         k = PG_GETARG_INT32(1);
    dnagpu_stats st;
    TupleDesc   tupdesc;
    Datum       values[3];
    bool        nulls[3] = {false, false, false};
    int         rc;

    check_k(k);
    if (get_call_result_type(fcinfo, NULL, &tupdesc) != TYPEFUNC_COMPOSITE)
        ereport(ERROR, (errmsg("kmer_stats must be called in a context that accepts a record")));
    rc = dnagpu_count_kmers(gpu(), dna->bit_sequence, dna->length, k, NULL, &st, NULL);
    if (rc != DNAGPU_OK)
        ereport(ERROR, (errmsg("%s", dnagpu_last_error(backend_ctx))));
    values[0] = Int64GetDatum((int64) st.total);
    values[1] = Int64GetDatum((int64) st.distinct);
    values[2] = Int64GetDatum((int64) st.unique);
    PG_FREE_IF_COPY(dna, 0);
    PG_RETURN_DATUM(HeapTupleGetDatum(heap_form_tuple(BlessTupleDesc(tupdesc), values, nulls)));
}

/*
 * count_kmers(dna, int) SETOF (kmer, count): SELECT kmer, count(*) FROM generate_kmers(...) GROUP BY kmer
 * (README.md:107-116, test.sql:95-104) as one function.  The first call groups on the GPU and fetches the
 * rows; the per-row calls only form tuples.  Row order is unspecified, like a HashAggregate's.
 */
typedef struct CountState
{
    uint64_t   *kmers;
    uint64_t   *counts;
    int         k;
} CountState;

PG_FUNCTION_INFO_V1(count_kmers);
Datum
count_kmers(PG_FUNCTION_ARGS)
{
    FuncCallContext *funcctx;
    CountState *state;

    if (SRF_IS_FIRSTCALL())
    {
        MemoryContext oldcontext;
        Dna        *dna;
        int         k, rc;
        uint64_t    rows;
        dnagpu_stats st;
        dnagpu_table *table = NULL;
        TupleDesc   tupdesc;

        funcctx = SRF_FIRSTCALL_INIT();
        oldcontext = MemoryContextSwitchTo(funcctx->multi_call_memory_ctx);
        dna = (Dna *) PG_GETARG_VARLENA_P(0);
        k = PG_GETARG_INT32(1);
        check_k(k);
        if (get_call_result_type(fcinfo, NULL, &tupdesc) != TYPEFUNC_COMPOSITE)
            ereport(ERROR, (errmsg("count_kmers must be called in a context that accepts a record")));
        funcctx->tuple_desc = BlessTupleDesc(tupdesc);

        rc = dnagpu_count_kmers(gpu(), dna->bit_sequence, dna->length, k, NULL, &st, &table);
        if (rc != DNAGPU_OK)
            ereport(ERROR, (errmsg("%s", dnagpu_last_error(backend_ctx))));
        rows = dnagpu_table_rows(table);
        state = (CountState *) palloc(sizeof(CountState));
        state->k = k;
        state->kmers = (uint64_t *) palloc(sizeof(uint64_t) * (rows ? rows : 1));
        state->counts = (uint64_t *) palloc(sizeof(uint64_t) * (rows ? rows : 1));
        rc = dnagpu_table_fetch(backend_ctx, table, 0, rows, state->kmers, state->counts);
        dnagpu_table_free(table);   /* before any ereport: the table is not in a memory context */
        if (rc != DNAGPU_OK)
            ereport(ERROR, (errmsg("%s", dnagpu_last_error(backend_ctx))));
        funcctx->user_fctx = state;
        funcctx->max_calls = rows;
        MemoryContextSwitchTo(oldcontext);
    }

    funcctx = SRF_PERCALL_SETUP();
    state = (CountState *) funcctx->user_fctx;

    if (funcctx->call_cntr < funcctx->max_calls)
    {
        Kmer       *kmer = (Kmer *) palloc0(sizeof(Kmer));
        Datum       values[2];
        bool        nulls[2] = {false, false};

        kmer->length = state->k;
        kmer->bit_sequence = state->kmers[funcctx->call_cntr];
        values[0] = PointerGetDatum(kmer);
        values[1] = Int64GetDatum((int64) state->counts[funcctx->call_cntr]);
        SRF_RETURN_NEXT(funcctx, HeapTupleGetDatum(heap_form_tuple((TupleDesc) funcctx->tuple_desc, values, nulls)));
    }
    SRF_RETURN_DONE(funcctx);
}
