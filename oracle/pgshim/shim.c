/*
 * pgshim/shim.c -- the few server routines dna.c links against, re-implemented for a
 * stand-alone build (see postgres.h).  TEST INFRASTRUCTURE ONLY.
 */
#include "postgres.h"

int shim_module_magic = 16;
__thread MemoryContext CurrentMemoryContext = NULL;

/* ---- memory: plain malloc; the driver frees what the executor would reset ---- */
void *palloc(Size size) { return malloc(size ? size : 1); }
void *palloc0(Size size) { return calloc(1, size ? size : 1); }
void *repalloc(void *p, Size size) { return realloc(p, size ? size : 1); }
void pfree(void *p) { free(p); }
char *pstrdup(const char *s)
{
    size_t n = strlen(s) + 1;
    char *d = (char *)malloc(n);
    memcpy(d, s, n);
    return d;
}

/* ---- ereport(ERROR) = longjmp to whoever called into dna.c ---- */
__thread char shim_error_text[512];
__thread jmp_buf *shim_error_jmp = NULL;

int shim_errmsg(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(shim_error_text, sizeof shim_error_text, fmt, ap);
    va_end(ap);
    return 0;
}

void shim_ereport(int level)
{
    if (level < ERROR) return; /* INFO / NOTICE / WARNING: message only */
    if (shim_error_jmp) longjmp(*shim_error_jmp, 1);
    fprintf(stderr, "pgshim: ERROR outside a guarded call: %s\n", shim_error_text);
    abort();
}

/* ---- text ---- */
text *cstring_to_text(const char *s)
{
    size_t n = strlen(s);
    text *t = (text *)palloc(n + VARHDRSZ);
    SET_VARSIZE(t, n + VARHDRSZ);
    memcpy(VARDATA(t), s, n);
    return t;
}
char *text_to_cstring(const text *t)
{
    size_t n = VARSIZE(t) - VARHDRSZ;
    char *s = (char *)palloc(n + 1);
    memcpy(s, VARDATA(t), n);
    s[n] = '\0';
    return s;
}

Datum textin(PG_FUNCTION_ARGS) { PG_RETURN_POINTER(cstring_to_text(PG_GETARG_CSTRING(0))); }
Datum textout(PG_FUNCTION_ARGS) { PG_RETURN_CSTRING(text_to_cstring((text *)PG_GETARG_POINTER(0))); }
Datum shim_DirectFunctionCall1(Datum (*fn)(PG_FUNCTION_ARGS), Datum arg)
{
    FunctionCallInfoBaseData fc;
    FmgrInfo fl;
    memset(&fc, 0, sizeof fc);
    memset(&fl, 0, sizeof fl);
    fc.flinfo = &fl;
    fc.nargs = 1;
    fc.args[0].value = arg;
    return fn(&fc);
}

/* ---- memory contexts: only their reset callbacks are modelled ---- */
struct MemoryContextData {
    MemoryContextCallback *callbacks;
    struct MemoryContextData *prev, *next;
};
static __thread struct MemoryContextData *live_contexts = NULL;

MemoryContext shim_context_create(void)
{
    MemoryContext c = (MemoryContext)calloc(1, sizeof(*c));
    c->next = live_contexts;
    if (live_contexts) live_contexts->prev = c;
    live_contexts = c;
    return c;
}
void shim_context_delete(MemoryContext c)
{
    if (!c) return;
    while (c->callbacks) { /* newest first, each exactly once */
        MemoryContextCallback *cb = c->callbacks;
        c->callbacks = cb->next;
        cb->func(cb->arg);
    }
    if (c->prev) c->prev->next = c->next; else live_contexts = c->next;
    if (c->next) c->next->prev = c->prev;
    free(c);
}
void shim_abort_cleanup(void)
{
    while (live_contexts) shim_context_delete(live_contexts);
}
int shim_live_contexts(void)
{
    int n = 0;
    struct MemoryContextData *c;
    for (c = live_contexts; c; c = c->next) n++;
    return n;
}
void MemoryContextRegisterResetCallback(MemoryContext context, MemoryContextCallback *cb)
{
    cb->next = context->callbacks;
    context->callbacks = cb;
}
void *MemoryContextAllocHuge(MemoryContext context, Size size)
{
    (void)context;
    return malloc(size ? size : 1);
}
void *repalloc_huge(void *p, Size size) { return realloc(p, size ? size : 1); }

int AggCheckCallContext(FunctionCallInfo fcinfo, MemoryContext *aggcontext)
{
    ShimAggContext *a = (ShimAggContext *)fcinfo->context;
    if (!a || a->tag != 0x4147) return 0;
    if (aggcontext) *aggcontext = a->aggcontext;
    return AGG_CONTEXT_AGGREGATE;
}

/* ---- SRF ---- */
FuncCallContext *shim_init_MultiFuncCall(FunctionCallInfo fcinfo)
{
    FuncCallContext *f = (FuncCallContext *)calloc(1, sizeof(*f));
    f->multi_call_memory_ctx = shim_context_create();
    fcinfo->flinfo->fn_extra = f;
    return f;
}
void shim_end_MultiFuncCall(FunctionCallInfo fcinfo, FuncCallContext *funcctx)
{
    shim_context_delete(funcctx->multi_call_memory_ctx); /* end_MultiFuncCall deletes the context: callbacks fire */
    free(funcctx);
    fcinfo->flinfo->fn_extra = NULL;
}

/* ---- composite results ---- */
TypeFuncClass get_call_result_type(FunctionCallInfo fcinfo, Oid *resultTypeId, TupleDesc *resultTupleDesc)
{
    if (resultTypeId) *resultTypeId = InvalidOid;
    if (!fcinfo->shim_result_desc) return TYPEFUNC_SCALAR;
    if (resultTupleDesc) *resultTupleDesc = (TupleDesc)fcinfo->shim_result_desc;
    return TYPEFUNC_COMPOSITE;
}
TupleDesc BlessTupleDesc(TupleDesc tupdesc) { return tupdesc; }
HeapTuple heap_form_tuple(TupleDesc tupdesc, const Datum *values, const bool *isnull)
{
    HeapTuple t = (HeapTuple)palloc(sizeof(HeapTupleData));
    t->natts = tupdesc->natts;
    t->values = (Datum *)palloc(sizeof(Datum) * (size_t)t->natts);
    t->nulls = (bool *)palloc(sizeof(bool) * (size_t)t->natts);
    memcpy(t->values, values, sizeof(Datum) * (size_t)t->natts);
    memcpy(t->nulls, isnull, sizeof(bool) * (size_t)t->natts);
    return t;
}
void heap_freetuple(HeapTuple t)
{
    pfree(t->values);
    pfree(t->nulls);
    pfree(t);
}

/* ---- pqformat (network byte order, like the server) ---- */
static void sb_need(StringInfo b, int n)
{
    if (b->len + n + 1 > b->maxlen) {
        b->maxlen = (b->len + n + 1) * 2;
        b->data = (char *)realloc(b->data, (size_t)b->maxlen);
    }
}
void pq_begintypsend(StringInfo buf)
{
    buf->maxlen = 64;
    buf->data = (char *)malloc(64);
    buf->len = 4; /* room for the varlena header */
    buf->cursor = 0;
}
bytea *pq_endtypsend(StringInfo buf)
{
    bytea *r = (bytea *)buf->data;
    SET_VARSIZE(r, buf->len);
    return r;
}
void pq_sendbytes(StringInfo buf, const void *data, int datalen)
{
    sb_need(buf, datalen);
    memcpy(buf->data + buf->len, data, (size_t)datalen);
    buf->len += datalen;
}
void pq_sendint(StringInfo buf, uint32 i, int b)
{
    unsigned char n[4];
    switch (b) {
    case 1: n[0] = (unsigned char)i; pq_sendbytes(buf, n, 1); break;
    case 2: n[0] = (unsigned char)(i >> 8); n[1] = (unsigned char)i; pq_sendbytes(buf, n, 2); break;
    case 4:
        n[0] = (unsigned char)(i >> 24); n[1] = (unsigned char)(i >> 16);
        n[2] = (unsigned char)(i >> 8); n[3] = (unsigned char)i;
        pq_sendbytes(buf, n, 4);
        break;
    default: /* the server: elog(ERROR, "unsupported integer size %d") -- SURVEY Q8 */
        shim_errmsg("unsupported integer size %d", b);
        shim_ereport(ERROR);
    }
}
void pq_sendint64(StringInfo buf, uint64 i)
{
    pq_sendint(buf, (uint32)(i >> 32), 4);
    pq_sendint(buf, (uint32)i, 4);
}
void pq_copymsgbytes(StringInfo msg, char *buf, int datalen)
{
    if (datalen < 0 || datalen > msg->len - msg->cursor) {
        shim_errmsg("insufficient data left in message");
        shim_ereport(ERROR);
    }
    memcpy(buf, msg->data + msg->cursor, (size_t)datalen);
    msg->cursor += datalen;
}
unsigned int pq_getmsgint(StringInfo msg, int b)
{
    unsigned char n[4];
    switch (b) {
    case 1: pq_copymsgbytes(msg, (char *)n, 1); return n[0];
    case 2: pq_copymsgbytes(msg, (char *)n, 2); return ((unsigned)n[0] << 8) | n[1];
    case 4:
        pq_copymsgbytes(msg, (char *)n, 4);
        return ((unsigned)n[0] << 24) | ((unsigned)n[1] << 16) | ((unsigned)n[2] << 8) | n[3];
    default:
        shim_errmsg("unsupported integer size %d", b);
        shim_ereport(ERROR);
    }
    return 0;
}
int64 pq_getmsgint64(StringInfo msg)
{
    uint64 hi = pq_getmsgint(msg, 4), lo = pq_getmsgint(msg, 4);
    return (int64)((hi << 32) | lo);
}

/* ---- hash_any: lookup3 (Bob Jenkins, public domain) with PostgreSQL's initial value and
 * its little-endian tail handling; byte-wise form, valid for any alignment ---- */
#define rot(x, k) (((x) << (k)) | ((x) >> (32 - (k))))
#define mix(a, b, c)                    \
    {                                   \
        a -= c; a ^= rot(c, 4);  c += b; \
        b -= a; b ^= rot(a, 6);  a += c; \
        c -= b; c ^= rot(b, 8);  b += a; \
        a -= c; a ^= rot(c, 16); c += b; \
        b -= a; b ^= rot(a, 19); a += c; \
        c -= b; c ^= rot(b, 4);  b += a; \
    }
#define final(a, b, c)        \
    {                         \
        c ^= b; c -= rot(b, 14); \
        a ^= c; a -= rot(c, 11); \
        b ^= a; b -= rot(a, 25); \
        c ^= b; c -= rot(b, 16); \
        a ^= c; a -= rot(c, 4);  \
        b ^= a; b -= rot(a, 14); \
        c ^= b; c -= rot(b, 24); \
    }
uint32 hash_bytes(const unsigned char *k, int keylen)
{
    uint32 a, b, c, len = (uint32)keylen;
    a = b = c = 0x9e3779b9 + len + 3923095;
    while (len >= 12) {
        a += k[0] + ((uint32)k[1] << 8) + ((uint32)k[2] << 16) + ((uint32)k[3] << 24);
        b += k[4] + ((uint32)k[5] << 8) + ((uint32)k[6] << 16) + ((uint32)k[7] << 24);
        c += k[8] + ((uint32)k[9] << 8) + ((uint32)k[10] << 16) + ((uint32)k[11] << 24);
        mix(a, b, c);
        k += 12;
        len -= 12;
    }
    switch (len) { /* the lowest byte of c is reserved for the length */
    case 11: c += ((uint32)k[10] << 24); /* fall through */
    case 10: c += ((uint32)k[9] << 16);  /* fall through */
    case 9: c += ((uint32)k[8] << 8);    /* fall through */
    case 8: b += ((uint32)k[7] << 24);   /* fall through */
    case 7: b += ((uint32)k[6] << 16);   /* fall through */
    case 6: b += ((uint32)k[5] << 8);    /* fall through */
    case 5: b += k[4];                   /* fall through */
    case 4: a += ((uint32)k[3] << 24);   /* fall through */
    case 3: a += ((uint32)k[2] << 16);   /* fall through */
    case 2: a += ((uint32)k[1] << 8);    /* fall through */
    case 1: a += k[0];
    }
    final(a, b, c);
    return c;
}

/* ---- catalog lookups used only by get_oid / spgist_kmer_config ---- */
TypeName *makeTypeName(char *typnam)
{
    TypeName *t = (TypeName *)palloc(sizeof(*t));
    t->name = typnam;
    return t;
}
Oid typenameTypeId(void *pstate, const TypeName *typeName)
{
    (void)pstate;
    (void)typeName;
    return 16385; /* some user-type OID */
}
