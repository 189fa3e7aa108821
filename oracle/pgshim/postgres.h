/*
 * pgshim/postgres.h -- a minimal stand-in for the PostgreSQL 16 server headers, just large
 * enough to compile the reference's UNMODIFIED dna.c outside a server (there is no
 * PostgreSQL in this image: no pg_config, no server headers).  TEST INFRASTRUCTURE ONLY.
 *
 * What is emulated: palloc/pfree (malloc), ereport(ERROR) (longjmp to the driver), varlena
 * headers, the fmgr V1 calling convention, the ValuePerCall SRF protocol of funcapi.h,
 * composite results (get_call_result_type / heap_form_tuple, for the GPU glue's pushdown
 * functions: dna.c itself returns none), pqformat buffers, hash_any (Bob Jenkins' lookup3 as PostgreSQL uses it) and the struct
 * layouts of access/spgist.h that dna.c's index support functions mention (never called).
 * Every other server header dna.c includes is an empty file that includes this one.
 */
#ifndef PGSHIM_POSTGRES_H
#define PGSHIM_POSTGRES_H

#include <setjmp.h>
#include <stdarg.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int16_t int16;
typedef int32_t int32;
typedef int64_t int64;
typedef uint8_t uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef uint64_t uint64;
typedef size_t Size;
typedef unsigned int Oid;
typedef uintptr_t Datum;
typedef char *Pointer;

#define FLEXIBLE_ARRAY_MEMBER
#define Min(x, y) ((x) < (y) ? (x) : (y))
#define Max(x, y) ((x) > (y) ? (x) : (y))
#define Assert(c) ((void)0) /* --enable-cassert is off in production builds */
#define PGDLLEXPORT
#define INT2OID 21
#define InvalidOid 0

/* ---- memory ---- */
void *palloc(Size size);
void *palloc0(Size size);
void *repalloc(void *p, Size size);
void pfree(void *p);
char *pstrdup(const char *s);
typedef struct MemoryContextData *MemoryContext;
extern __thread MemoryContext CurrentMemoryContext;
/* utils/memutils.h, utils/palloc.h: contexts exist here only to carry reset callbacks (what lets an extension
 * release non-palloc resources when a query ends or aborts); allocations stay plain malloc */
typedef void (*MemoryContextCallbackFunction)(void *arg);
typedef struct MemoryContextCallback {
    MemoryContextCallbackFunction func;
    void *arg;
    struct MemoryContextCallback *next;
} MemoryContextCallback;
void MemoryContextRegisterResetCallback(MemoryContext context, MemoryContextCallback *cb);
void *MemoryContextAllocHuge(MemoryContext context, Size size);
void *repalloc_huge(void *p, Size size);
MemoryContext shim_context_create(void);
void shim_context_delete(MemoryContext c); /* fires the callbacks, newest first, like MemoryContextDelete */
void shim_abort_cleanup(void);             /* transaction abort: every live context is deleted */
int shim_live_contexts(void);
static inline MemoryContext MemoryContextSwitchTo(MemoryContext c)
{
    MemoryContext old = CurrentMemoryContext;
    CurrentMemoryContext = c;
    return old;
}

/* ---- errors ---- */
#define DEBUG1 14
#define INFO 17
#define NOTICE 18
#define WARNING 19
#define ERROR 21
int shim_errmsg(const char *fmt, ...);
void shim_ereport(int level);
#define errmsg shim_errmsg
#define ereport(level, rest) do { (void)(rest); shim_ereport(level); } while (0)
#define elog(level, ...) do { shim_errmsg(__VA_ARGS__); shim_ereport(level); } while (0)

/* ---- varlena ---- */
struct varlena {
    char vl_len_[4];
    char vl_dat[FLEXIBLE_ARRAY_MEMBER];
};
typedef struct varlena text;
typedef struct varlena bytea;
#define VARHDRSZ ((int32)sizeof(int32))
#define SET_VARSIZE(p, len) (*(uint32 *)(p) = (uint32)(len) << 2) /* 4-byte header, little-endian build */
#define VARSIZE(p) ((*(uint32 *)(p)) >> 2)
#define VARDATA(p) (((char *)(p)) + VARHDRSZ)
#define VARSIZE_ANY_EXHDR(p) (VARSIZE(p) - VARHDRSZ)
#define VARDATA_ANY(p) VARDATA(p)
text *cstring_to_text(const char *s);
char *text_to_cstring(const text *t);

/* ---- Datum ---- */
#define PointerGetDatum(x) ((Datum)(uintptr_t)(x))
#define DatumGetPointer(x) ((Pointer)(uintptr_t)(x))
#define CStringGetDatum(x) PointerGetDatum(x)
#define DatumGetCString(x) ((char *)DatumGetPointer(x))
#define Int16GetDatum(x) ((Datum)(uintptr_t)(int16)(x))
#define DatumGetInt16(x) ((int16)(x))
#define Int32GetDatum(x) ((Datum)(uintptr_t)(int32)(x))
#define DatumGetInt32(x) ((int32)(x))
#define UInt32GetDatum(x) ((Datum)(uintptr_t)(uint32)(x))
#define DatumGetUInt32(x) ((uint32)(x))
#define Int64GetDatum(x) ((Datum)(int64)(x)) /* 64-bit build: pass-by-value */
#define DatumGetInt64(x) ((int64)(x))
#define BoolGetDatum(x) ((Datum)((x) ? 1 : 0))
#define DatumGetBool(x) ((bool)((x) != 0))

/* ---- fmgr V1 ---- */
typedef struct FmgrInfo {
    void *fn_extra;
    MemoryContext fn_mcxt;
} FmgrInfo;
typedef struct NullableDatum {
    Datum value;
    bool isnull;
} NullableDatum;
typedef struct FunctionCallInfoBaseData {
    FmgrInfo *flinfo;
    void *context;
    void *resultinfo;
    Oid fncollation;
    bool isnull;
    short nargs;
    NullableDatum args[8];
    void *shim_result_desc; /* shim only: the row type the caller expects (the catalog lookup of the server) */
} FunctionCallInfoBaseData;
typedef FunctionCallInfoBaseData *FunctionCallInfo;
#define PG_FUNCTION_ARGS FunctionCallInfo fcinfo
#define PG_MODULE_MAGIC extern int shim_module_magic
#define PG_FUNCTION_INFO_V1(f) extern Datum f(PG_FUNCTION_ARGS)
#define PG_NARGS() (fcinfo->nargs)
#define PG_ARGISNULL(n) (fcinfo->args[n].isnull)
#define PG_GETARG_DATUM(n) (fcinfo->args[n].value)
#define PG_GETARG_POINTER(n) DatumGetPointer(PG_GETARG_DATUM(n))
#define PG_GETARG_CSTRING(n) DatumGetCString(PG_GETARG_DATUM(n))
#define PG_GETARG_INT32(n) DatumGetInt32(PG_GETARG_DATUM(n))
#define PG_GETARG_VARLENA_P(n) ((struct varlena *)PG_GETARG_POINTER(n)) /* nothing is ever toasted here */
#define PG_GETARG_TEXT_P(n) ((text *)PG_GETARG_POINTER(n))
#define PG_FREE_IF_COPY(p, n) ((void)0)
#define PG_RETURN_DATUM(x) return (x)
#define PG_RETURN_POINTER(x) return PointerGetDatum(x)
#define PG_RETURN_CSTRING(x) return CStringGetDatum(x)
#define PG_RETURN_TEXT_P(x) PG_RETURN_POINTER(x)
#define PG_RETURN_BYTEA_P(x) PG_RETURN_POINTER(x)
#define PG_RETURN_BOOL(x) return BoolGetDatum(x)
#define PG_RETURN_INT32(x) return Int32GetDatum(x)
#define PG_RETURN_UINT32(x) return UInt32GetDatum(x)
#define PG_RETURN_VOID() return (Datum)0
#define PG_RETURN_NULL() do { fcinfo->isnull = true; return (Datum)0; } while (0)

Datum textin(PG_FUNCTION_ARGS);
Datum textout(PG_FUNCTION_ARGS);
Datum shim_DirectFunctionCall1(Datum (*fn)(PG_FUNCTION_ARGS), Datum arg);
#define DirectFunctionCall1(fn, arg) shim_DirectFunctionCall1(fn, arg)

/* ---- funcapi.h: ValuePerCall set-returning functions ---- */
typedef enum { ExprSingleResult, ExprMultipleResult, ExprEndResult } ExprDoneCond;
typedef struct ReturnSetInfo {
    ExprDoneCond isDone;
} ReturnSetInfo;
typedef struct FuncCallContext {
    uint64 call_cntr;
    uint64 max_calls;
    void *user_fctx;
    void *attinmeta;
    MemoryContext multi_call_memory_ctx;
    void *tuple_desc;
} FuncCallContext;
FuncCallContext *shim_init_MultiFuncCall(FunctionCallInfo fcinfo);
#define SRF_IS_FIRSTCALL() (fcinfo->flinfo->fn_extra == NULL)
#define SRF_FIRSTCALL_INIT() shim_init_MultiFuncCall(fcinfo)
#define SRF_PERCALL_SETUP() ((FuncCallContext *)fcinfo->flinfo->fn_extra)
#define SRF_RETURN_NEXT(_funcctx, _result)                        \
    do {                                                          \
        ReturnSetInfo *rsi;                                       \
        (_funcctx)->call_cntr++;                                  \
        rsi = (ReturnSetInfo *)fcinfo->resultinfo;                \
        rsi->isDone = ExprMultipleResult;                         \
        PG_RETURN_DATUM(_result);                                 \
    } while (0)
void shim_end_MultiFuncCall(FunctionCallInfo fcinfo, FuncCallContext *funcctx);
#define SRF_RETURN_DONE(_funcctx)                                 \
    do {                                                          \
        ReturnSetInfo *rsi;                                       \
        shim_end_MultiFuncCall(fcinfo, _funcctx);                 \
        rsi = (ReturnSetInfo *)fcinfo->resultinfo;                \
        rsi->isDone = ExprEndResult;                              \
        PG_RETURN_NULL();                                         \
    } while (0)

/* ---- fmgr.h: aggregate support functions ---- */
#define AGG_CONTEXT_AGGREGATE 1
typedef struct ShimAggContext { /* what fcinfo->context points at when the executor calls an sfunc / finalfunc */
    int tag;                    /* 0x4147 */
    MemoryContext aggcontext;
} ShimAggContext;
int AggCheckCallContext(FunctionCallInfo fcinfo, MemoryContext *aggcontext);

/* ---- funcapi.h / access/htup_details.h: composite results ---- */
typedef struct TupleDescData {
    int natts;
} TupleDescData;
typedef TupleDescData *TupleDesc;
typedef enum { TYPEFUNC_SCALAR, TYPEFUNC_COMPOSITE, TYPEFUNC_COMPOSITE_DOMAIN, TYPEFUNC_RECORD, TYPEFUNC_OTHER } TypeFuncClass;
TypeFuncClass get_call_result_type(FunctionCallInfo fcinfo, Oid *resultTypeId, TupleDesc *resultTupleDesc);
TupleDesc BlessTupleDesc(TupleDesc tupdesc);
typedef struct HeapTupleData { /* the shim keeps the column values as they are */
    int natts;
    Datum *values;
    bool *nulls;
} HeapTupleData;
typedef HeapTupleData *HeapTuple;
HeapTuple heap_form_tuple(TupleDesc tupdesc, const Datum *values, const bool *isnull);
void heap_freetuple(HeapTuple t);
#define HeapTupleGetDatum(t) PointerGetDatum(t)

/* ---- libpq/pqformat.h ---- */
typedef struct StringInfoData {
    char *data;
    int len;
    int maxlen;
    int cursor;
} StringInfoData;
typedef StringInfoData *StringInfo;
void pq_begintypsend(StringInfo buf);
bytea *pq_endtypsend(StringInfo buf);
void pq_sendint(StringInfo buf, uint32 i, int b);
void pq_sendint64(StringInfo buf, uint64 i);
void pq_sendbytes(StringInfo buf, const void *data, int datalen);
unsigned int pq_getmsgint(StringInfo msg, int b);
int64 pq_getmsgint64(StringInfo msg);
void pq_copymsgbytes(StringInfo msg, char *buf, int datalen);

/* ---- common/hashfn.h ---- */
uint32 hash_bytes(const unsigned char *k, int keylen);
static inline Datum hash_any(const unsigned char *k, int keylen) { return UInt32GetDatum(hash_bytes(k, keylen)); }

/* ---- parser/parse_type.h, nodes/makefuncs.h (only get_oid / spgist config use them) ---- */
typedef struct TypeName {
    const char *name;
} TypeName;
TypeName *makeTypeName(char *typnam);
Oid typenameTypeId(void *pstate, const TypeName *typeName);

/* ---- access/spgist.h: struct layouts only ---- */
typedef uint16 StrategyNumber;
#define BTEqualStrategyNumber 3
#define RTPrefixStrategyNumber 28
#define RTContainsStrategyNumber 7
#define RTContainedByStrategyNumber 8
typedef struct ScanKeyData {
    int sk_flags;
    int16 sk_attno;
    StrategyNumber sk_strategy;
    Oid sk_subtype;
    Oid sk_collation;
    FmgrInfo sk_func;
    Datum sk_argument;
} ScanKeyData;
typedef ScanKeyData *ScanKey;
typedef struct spgConfigIn {
    Oid attType;
} spgConfigIn;
typedef struct spgConfigOut {
    Oid prefixType;
    Oid labelType;
    Oid leafType;
    bool canReturnData;
    bool longValuesOK;
} spgConfigOut;
typedef struct spgChooseIn {
    Datum datum;
    Datum leafDatum;
    int level;
    bool allTheSame;
    bool hasPrefix;
    Datum prefixDatum;
    int nNodes;
    Datum *nodeLabels;
} spgChooseIn;
typedef enum spgChooseResultType { spgMatchNode = 1, spgAddNode, spgSplitTuple } spgChooseResultType;
typedef struct spgChooseOut {
    spgChooseResultType resultType;
    union {
        struct {
            int nodeN;
            int levelAdd;
            Datum restDatum;
        } matchNode;
        struct {
            Datum nodeLabel;
            int nodeN;
        } addNode;
        struct {
            bool prefixHasPrefix;
            Datum prefixPrefixDatum;
            int prefixNNodes;
            Datum *prefixNodeLabels;
            int childNodeN;
            bool postfixHasPrefix;
            Datum postfixPrefixDatum;
        } splitTuple;
    } result;
} spgChooseOut;
typedef struct spgPickSplitIn {
    int nTuples;
    Datum *datums;
    int level;
} spgPickSplitIn;
typedef struct spgPickSplitOut {
    bool hasPrefix;
    Datum prefixDatum;
    int nNodes;
    Datum *nodeLabels;
    int *mapTuplesToNodes;
    Datum *leafTupleDatums;
} spgPickSplitOut;
typedef struct spgInnerConsistentIn {
    ScanKey scankeys;
    ScanKey orderbys;
    int nkeys;
    int norderbys;
    Datum reconstructedValue;
    void *traversalValue;
    MemoryContext traversalMemoryContext;
    int level;
    bool returnData;
    bool allTheSame;
    bool hasPrefix;
    Datum prefixDatum;
    int nNodes;
    Datum *nodeLabels;
} spgInnerConsistentIn;
typedef struct spgInnerConsistentOut {
    int nNodes;
    int *nodeNumbers;
    int *levelAdds;
    Datum *reconstructedValues;
    void **traversalValues;
    double **distances;
} spgInnerConsistentOut;
typedef struct spgLeafConsistentIn {
    ScanKey scankeys;
    ScanKey orderbys;
    int nkeys;
    int norderbys;
    Datum reconstructedValue;
    void *traversalValue;
    int level;
    bool returnData;
    Datum leafDatum;
} spgLeafConsistentIn;
typedef struct spgLeafConsistentOut {
    Datum leafValue;
    bool recheck;
    bool recheckDistances;
    double *distances;
} spgLeafConsistentOut;

#endif
