/*
 * index.cuh -- kernels of the sorted k-mer index: the replacement for the reference's SP-GiST
 * trie over a kmer column (spgist_kmer_ops, dna.c:1137-1737; used by test.sql:186-262 for
 * `kmer = x` and `kmer ^@ prefix`).
 *
 * The sort key of a k-mer is its base string read as a number with base 0 most significant
 * (the 2-bit groups of Kmer.bit_sequence in reverse order, dna.c:397-420 stores base j at bits
 * 2j..2j+1): all k-mers that start with a given prefix are then ONE contiguous range of the
 * sorted column, and an equality or prefix search is two binary searches.
 *
 * Build = stable LSD radix sort, 8 bits per pass, of (sort key, row number) pairs.
 * Every pass: per-CTA digit histogram -> exclusive scan in digit-major order -> stable scatter
 * (ranks from warp ballots, round-major inside a warp, warps and CTAs in input order).
 */
#pragma once
#include "kernels.cuh"

namespace dnagpu {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortChunk = kSortThreads * kSortItems; /* keys per CTA and chunk */
constexpr int kSortWarps = kSortThreads / 32;

/* reverse the order of the 32 two-bit groups of x */
__host__ __device__ __forceinline__ uint64_t rev_pairs(uint64_t x)
{
#ifdef __CUDA_ARCH__
    x = __brevll(x);
#else
    x = ((x >> 32) | (x << 32));
    x = ((x & 0xffff0000ffff0000ull) >> 16) | ((x & 0x0000ffff0000ffffull) << 16);
    x = ((x & 0xff00ff00ff00ff00ull) >> 8) | ((x & 0x00ff00ff00ff00ffull) << 8);
    x = ((x & 0xf0f0f0f0f0f0f0f0ull) >> 4) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
    x = ((x & 0xccccccccccccccccull) >> 2) | ((x & 0x3333333333333333ull) << 2);
    x = ((x & 0xaaaaaaaaaaaaaaaaull) >> 1) | ((x & 0x5555555555555555ull) << 1);
#endif
    /* a full bit reversal also swapped the two bits of every base: swap them back */
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}
/* Kmer.bit_sequence (k bases) <-> sort key (base 0 in the top two of 2k bits) */
__host__ __device__ __forceinline__ uint64_t sort_key_of(uint64_t kmer, int k) { return rev_pairs(kmer) >> (64 - 2 * k); }
__host__ __device__ __forceinline__ uint64_t kmer_of_sort_key(uint64_t s, int k) { return rev_pairs(s << (64 - 2 * k)); }

__global__ void __launch_bounds__(kThreads) k_index_keys(const uint64_t *__restrict__ kmers, uint64_t n, int k,
                                                         uint64_t *__restrict__ skeys, uint64_t *__restrict__ rows)
{
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads) {
        skeys[i] = sort_key_of(ld_nc(kmers + i), k);
        rows[i] = i;
    }
}

/* cnt[d * gridDim.x + cta] = keys of digit d in the CTA's range [cta * per_cta, ...) */
__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint64_t *__restrict__ keys, uint64_t n,
                                                            uint64_t per_cta, int shift, uint64_t *__restrict__ cnt)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t beg = (uint64_t)blockIdx.x * per_cta, end = beg + per_cta < n ? beg + per_cta : n;
    for (uint64_t i = beg + threadIdx.x; i < end; i += kSortThreads)
        atomicAdd(&h[(uint32_t)(ld_nc(keys + i) >> shift) & 255u], 1u);
    __syncthreads();
    cnt[(uint64_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

/* Stable scatter of one pass.  off = exclusive scan of cnt (digit-major), i.e. where the CTA's first
 * key of every digit goes.  per_cta is a multiple of kSortChunk.
 * A chunk of 4096 pairs is ranked (warp ballots, round-major inside a warp, warps in input order), laid out by
 * digit in shared memory and copied out run by run: consecutive threads store consecutive addresses of a
 * digit's run instead of one 32-byte sector per lane straight from the registers. */
struct SortSmem {
    uint64_t keys[kSortChunk];
    uint64_t vals[kSortChunk];
    uint32_t wcnt[kSortWarps][256]; /* keys of digit d seen by warp w, then the warp's start inside the digit */
    uint64_t dbase[256];            /* where the next key of digit d goes in the output */
    uint32_t dstart[256 + 1];       /* start of digit d inside the staged chunk */
    uint32_t wtot[kSortWarps];
};

template <bool VALS>
__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint64_t *__restrict__ keys,
                                                               const uint64_t *__restrict__ vals, uint64_t n,
                                                               uint64_t per_cta, int shift,
                                                               const uint64_t *__restrict__ off,
                                                               uint64_t *__restrict__ keys_out,
                                                               uint64_t *__restrict__ vals_out)
{
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem &s = *reinterpret_cast<SortSmem *>(sort_smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    s.dbase[tid] = off[(uint64_t)tid * gridDim.x + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) s.wcnt[w][tid] = 0;
    __syncthreads();
    const uint64_t beg = (uint64_t)blockIdx.x * per_cta, end = beg + per_cta < n ? beg + per_cta : n;
    for (uint64_t chunk = beg; chunk < end; chunk += kSortChunk) {
        uint64_t x[kSortItems], v[kSortItems];
        uint32_t rk[kSortItems];
        const uint64_t wbase = chunk + (uint64_t)warp * (32 * kSortItems) + lane;
#pragma unroll
        for (int j = 0; j < kSortItems; ++j) {
            const uint64_t i = wbase + (uint64_t)j * 32;
            const bool ok = i < end;
            x[j] = ok ? ld_nc(keys + i) : 0;
            if (VALS) v[j] = ok ? ld_nc(vals + i) : 0;
            const uint32_t d = (uint32_t)(x[j] >> shift) & 255u;
            uint32_t peers = __ballot_sync(0xffffffffu, ok); /* lanes of this round with my digit */
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const uint32_t vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                peers &= ((d >> b) & 1u) ? vote : ~vote;
            }
            uint32_t base = 0;
            if (ok) base = s.wcnt[warp][d];
            __syncwarp();
            if (ok && (peers & lt) == 0) s.wcnt[warp][d] = base + __popc(peers); /* the lowest peer updates */
            __syncwarp();
            rk[j] = ok ? base + __popc(peers & lt) : 0xffffffffu;
        }
        __syncthreads();
        uint32_t run = 0; /* thread d: keys of digit d per warp -> exclusive start of every warp inside the digit */
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t t = s.wcnt[w][tid];
            s.wcnt[w][tid] = run;
            run += t;
        }
        /* exclusive scan of the 256 digit totals -> start of every digit inside the staged chunk */
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s.wtot[warp] = inc;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t t = s.wtot[w];
            if (w < (int)warp) woff += t;
            total += t;
        }
        s.dstart[tid] = woff + inc - run;
        if (tid == 0) s.dstart[256] = total;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kSortItems; ++j) {
            if (rk[j] == 0xffffffffu) continue;
            const uint32_t d = (uint32_t)(x[j] >> shift) & 255u;
            const uint32_t pos = s.dstart[d] + s.wcnt[warp][d] + rk[j];
            s.keys[pos] = x[j];
            if (VALS) s.vals[pos] = v[j];
        }
        __syncthreads();
        for (uint32_t i = tid; i < total; i += kSortThreads) {
            const uint64_t key = s.keys[i];
            const uint32_t d = (uint32_t)(key >> shift) & 255u;
            const uint64_t pos = s.dbase[d] + (i - s.dstart[d]);
            keys_out[pos] = key;
            if (VALS) vals_out[pos] = s.vals[i];
        }
        __syncthreads();
        s.dbase[tid] += run;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) s.wcnt[w][tid] = 0;
        __syncthreads();
    }
}

/* out[0] = first position with skeys >= lo, out[1] = first position with skeys > hi */
__global__ void k_index_bounds(const uint64_t *__restrict__ skeys, uint64_t n, uint64_t lo, uint64_t hi,
                               uint64_t *__restrict__ out)
{
    if (threadIdx.x > 1) return;
    const bool upper = threadIdx.x == 1;
    const uint64_t key = upper ? hi : lo;
    uint64_t a = 0, b = n;
    while (a < b) {
        const uint64_t m = a + ((b - a) >> 1);
        const uint64_t s = skeys[m];
        if (upper ? s <= key : s < key) a = m + 1;
        else b = m;
    }
    out[threadIdx.x] = a;
}

/* rows of [beg, end) whose k-mer passes the predicate, appended in no particular order */
__global__ void __launch_bounds__(kThreads) k_index_match(const uint64_t *__restrict__ skeys,
                                                          const uint64_t *__restrict__ rows, uint64_t beg,
                                                          uint64_t end, int k, Pred p, uint64_t cap,
                                                          unsigned long long *__restrict__ cursor,
                                                          uint64_t *__restrict__ out)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t span = end - beg, padded = (span + 31) & ~31ull;
    for (uint64_t t = (uint64_t)blockIdx.x * kThreads + threadIdx.x; t < padded; t += (uint64_t)gridDim.x * kThreads) {
        const bool in = t < span;
        const bool hit = in && pred_ok(p, kmer_of_sort_key(ld_nc(skeys + beg + t), k));
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (!m) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint64_t pos = base + __popc(m & ((1u << lane) - 1u));
        if (hit && pos < cap) out[pos] = ld_nc(rows + beg + t);
    }
}

} // namespace dnagpu
