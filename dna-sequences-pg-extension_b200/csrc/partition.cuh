/*
 * partition.cuh -- GROUP BY kmer for large inputs: radix partition + shared-memory count.
 *
 * Why: an open-addressing table in HBM costs one random 32-byte-sector atomic per k-mer;
 * ncu (profiles/r01a_*) shows 122 B of DRAM traffic per insert at 1.9 TB/s -- random
 * access, not arithmetic, is the bound (15 Gkmer/s on 100 Mbp, 9 Gkmer/s on 3.1 Gbp).
 * Shared-memory atomics, by contrast, run at > 1 T updates/s chip-wide (k <= 7 dense
 * kernel).  So: move the k-mers with STREAMING traffic until every bucket fits in
 * shared memory, then count there.
 *
 *   level 1   k-mers (from packed words, WHERE fused) or a key list -> P1 <= 2048
 *             partitions by the top bits of mix64(kmer): histogram pass (exact sizes,
 *             no slack, no overflow), then scatter: a CTA ranks a tile of 8192 keys by
 *             partition in shared memory and writes one contiguous run per partition.
 *   level 2   the same inside every partition on the next hash bits (P2 <= 2048), so
 *             that a bucket holds ~1-3 K keys.  Skipped when level 1 is already enough.
 *   count     one CTA per bucket (persistent grid): 4096-slot shared-memory table,
 *             atomicCAS to claim a slot (a claim IS the first occurrence; the counter
 *             holds only the extra ones), then distinct / unique / rows from a scan of
 *             the 48 KB table.  A key that finds the table full (more distinct keys in
 *             a bucket than slots) spills to a small HBM table -- same Slot / hash_insert
 *             as the plain hash variant -- so any key distribution stays correct.
 *
 * Algorithmic traffic: 8 B written + 8 B read per key and level, + 8 B read per hist2
 * pass and for the count = 40 B/key for two levels from packed input (56 B/key from a key
 * list), all of it coalesced.
 */
#pragma once
#include "kernels.cuh"

namespace dnagpu {

constexpr int kTileKeys = kThreads * 32;   /* 8192 keys staged per scatter tile (64 KB)   */
constexpr int kSuperTile = kTileKeys * 8;  /* keys per CTA in a histogram pass              */
constexpr int kMaxFan = 2048;              /* partitions per level                          */
constexpr int kBucketSlots = 4096;         /* shared-memory table of the count kernel       */
constexpr uint64_t kSlotMul = 0x9E3779B97F4A7C15ull;

__device__ __forceinline__ uint32_t digit_of(uint64_t h, int shift, uint32_t fan_mask)
{
    return (uint32_t)(h >> shift) & fan_mask;
}

/* ---- tiles of a partitioned key array ------------------------------------------- */
/* tiles[p] = ceil(size_p / tile_keys) */
__global__ void __launch_bounds__(kThreads) k_part_tiles(const uint64_t *__restrict__ parent_off,
                                                         uint64_t n_parents, uint64_t tile_keys,
                                                         uint64_t *__restrict__ tiles)
{
    uint64_t p = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (p < n_parents) tiles[p] = (parent_off[p + 1] - parent_off[p] + tile_keys - 1) / tile_keys;
}

/* CTA -> (parent partition, key range) */
__device__ __forceinline__ void tile_range(const uint64_t *parent_off, const uint64_t *tile_off,
                                           uint64_t n_parents, uint64_t tile_keys, uint64_t &parent,
                                           uint64_t &beg, uint64_t &end)
{
    parent = n_parents == 1 ? 0 : upper_seq(tile_off, n_parents, blockIdx.x);
    uint64_t t = blockIdx.x - tile_off[parent];
    beg = parent_off[parent] + t * tile_keys;
    end = min(beg + tile_keys, parent_off[parent + 1]);
}

/* ---- histogram passes ---------------------------------------------------------------- */
template <int L, bool FILTER>
__global__ void __launch_bounds__(kThreads) k_part_hist_seq(SeqView sv, Pred p, uint64_t mask, int shift,
                                                            uint32_t fan,
                                                            unsigned long long *__restrict__ hist)
{
    __shared__ uint32_t h[kMaxFan];
    for (uint32_t i = threadIdx.x; i < fan; i += kThreads) h[i] = 0;
    __syncthreads();
    const uint32_t fm = fan - 1;
    for (uint64_t t = (uint64_t)blockIdx.x * kThreads + threadIdx.x; t < sv.n_items;
         t += (uint64_t)gridDim.x * kThreads) {
        int c;
        uint64_t row0;
        const uint64_t *w = locate_item<L>(sv, t, c, row0);
        uint64_t w0 = ld_nc(w), w1 = ld_nc(w + 1);
        roll_item<8>(w0, w1, c, [&](uint64_t x, int) {
            if (FILTER && !pred_ok(p, x)) return;
            x &= mask;
            if (x == kEmpty) return; /* 'G' x 32: side counter, added by the scatter pass */
            atomicAdd(&h[digit_of(mix64(x), shift, fm)], 1u);
        });
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < fan; i += kThreads)
        if (h[i]) atomicAdd(&hist[i], (unsigned long long)h[i]);
}

/* children of every parent partition: hist[parent * fan + digit] */
__global__ void __launch_bounds__(kThreads) k_part_hist_keys(const uint64_t *__restrict__ keys,
                                                             const uint64_t *__restrict__ parent_off,
                                                             const uint64_t *__restrict__ tile_off,
                                                             uint64_t n_parents, int shift, uint32_t fan,
                                                             unsigned long long *__restrict__ hist)
{
    __shared__ uint32_t h[kMaxFan];
    if (blockIdx.x >= tile_off[n_parents]) return; /* the grid is an upper bound on the tiles */
    for (uint32_t i = threadIdx.x; i < fan; i += kThreads) h[i] = 0;
    __syncthreads();
    uint64_t parent, beg, end;
    tile_range(parent_off, tile_off, n_parents, kSuperTile, parent, beg, end);
    const uint32_t fm = fan - 1;
    for (uint64_t base = beg; base < end; base += kThreads * 8) {
        uint64_t x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint64_t i = base + (uint64_t)u * kThreads + threadIdx.x;
            x[u] = i < end ? ld_nc(keys + i) : kEmpty;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (x[u] != kEmpty) atomicAdd(&h[digit_of(mix64(x[u]), shift, fm)], 1u);
    }
    __syncthreads();
    unsigned long long *dst = hist + parent * fan;
    for (uint32_t i = threadIdx.x; i < fan; i += kThreads)
        if (h[i]) atomicAdd(&dst[i], (unsigned long long)h[i]);
}

/* ---- scatter passes ------------------------------------------------------------------------ */
/* Shared state of one scatter tile.  `cur[d]` first counts, then (after the scan) is the
 * running write position of digit d inside the stage; gdelta[d] turns a stage position into
 * the global output index. */
struct ScatterSmem {
    uint32_t cur[kMaxFan];
    long long gdelta[kMaxFan];
    uint16_t dig[kTileKeys];
    uint32_t warp_tot[kThreads / 32];
};

/* after counting: scan cur[], claim the global runs, leave cur[] = stage start per digit */
__device__ __forceinline__ uint32_t scatter_plan(ScatterSmem &s, uint32_t fan,
                                                 const uint64_t *__restrict__ child_off,
                                                 unsigned long long *__restrict__ child_cur)
{
    /* exclusive scan of cur[0..fan) with kThreads threads, fan <= 8 * kThreads */
    const int per = kMaxFan / kThreads; /* 8 */
    uint32_t v[per], sum = 0;
    const uint32_t base = threadIdx.x * per;
#pragma unroll
    for (int i = 0; i < per; ++i) {
        v[i] = base + i < fan ? s.cur[base + i] : 0;
        sum += v[i];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s.warp_tot[wid] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) {
        uint32_t t = s.warp_tot[i];
        if (i < wid) woff += t;
        total += t;
    }
    uint32_t ex = woff + inc - sum;
#pragma unroll
    for (int i = 0; i < per; ++i) {
        if (base + i < fan) {
            s.cur[base + i] = ex;
            if (v[i])
                s.gdelta[base + i] = (long long)(child_off[base + i] +
                                                 atomicAdd(&child_cur[base + i], (unsigned long long)v[i])) -
                                     (long long)ex;
        }
        ex += v[i];
    }
    __syncthreads();
    return total;
}

__device__ __forceinline__ void scatter_flush(const ScatterSmem &s, const uint64_t *stage, uint32_t total,
                                              uint64_t *__restrict__ out)
{
    for (uint32_t i = threadIdx.x; i < total; i += kThreads)
        out[(uint64_t)(s.gdelta[s.dig[i]] + (long long)i)] = stage[i];
}

template <int L, bool FILTER>
__global__ void __launch_bounds__(kThreads) k_part_scatter_seq(SeqView sv, Pred p, uint64_t mask, int shift,
                                                               uint32_t fan,
                                                               const uint64_t *__restrict__ child_off,
                                                               unsigned long long *__restrict__ child_cur,
                                                               uint64_t *__restrict__ out,
                                                               unsigned long long *__restrict__ ctr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *stage = reinterpret_cast<uint64_t *>(smem_raw);
    ScatterSmem &s = *reinterpret_cast<ScatterSmem *>(smem_raw + sizeof(uint64_t) * kTileKeys);
    for (uint32_t i = threadIdx.x; i < fan; i += kThreads) s.cur[i] = 0;
    __syncthreads();
    const uint32_t fm = fan - 1;
    uint64_t t = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    uint64_t w0 = 0, w1 = 0;
    int c = 0;
    uint32_t side = 0, kept = 0;
    if (t < sv.n_items) {
        uint64_t row0;
        const uint64_t *w = locate_item<L>(sv, t, c, row0);
        w0 = ld_nc(w);
        w1 = ld_nc(w + 1);
        roll_item<8>(w0, w1, c, [&](uint64_t x, int) {
            if (FILTER && !pred_ok(p, x)) return;
            kept++;
            x &= mask;
            if (x == kEmpty) {
                side++;
                return;
            }
            atomicAdd(&s.cur[digit_of(mix64(x), shift, fm)], 1u);
        });
    }
    __syncthreads();
    const uint32_t total = scatter_plan(s, fan, child_off, child_cur);
    roll_item<8>(w0, w1, c, [&](uint64_t x, int) {
        if (FILTER && !pred_ok(p, x)) return;
        x &= mask;
        if (x == kEmpty) return;
        uint32_t d = digit_of(mix64(x), shift, fm);
        uint32_t pos = atomicAdd(&s.cur[d], 1u);
        stage[pos] = x;
        s.dig[pos] = (uint16_t)d;
    });
    __syncthreads();
    scatter_flush(s, stage, total, out);
    kept = warp_sum32(kept);
    side = warp_sum32(side);
    if ((threadIdx.x & 31) == 0) {
        if (kept) atomicAdd(&ctr[C_TOTAL], (unsigned long long)kept);
        if (side) atomicAdd(&ctr[C_SIDE], (unsigned long long)side);
    }
}

/* scatter the keys of every parent partition into its children: out[child_off[parent*fan+d] ...].
 * COUNT_SIDE: first level over a raw key list (the caller's keys may hold 'G' x 32). */
template <bool COUNT_SIDE>
__global__ void __launch_bounds__(kThreads) k_part_scatter_keys(const uint64_t *__restrict__ keys,
                                                                const uint64_t *__restrict__ parent_off,
                                                                const uint64_t *__restrict__ tile_off,
                                                                uint64_t n_parents, int shift, uint32_t fan,
                                                                const uint64_t *__restrict__ child_off,
                                                                unsigned long long *__restrict__ child_cur,
                                                                uint64_t *__restrict__ out,
                                                                unsigned long long *__restrict__ ctr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *stage = reinterpret_cast<uint64_t *>(smem_raw);
    ScatterSmem &s = *reinterpret_cast<ScatterSmem *>(smem_raw + sizeof(uint64_t) * kTileKeys);
    if (blockIdx.x >= tile_off[n_parents]) return; /* the grid is an upper bound on the tiles */
    for (uint32_t i = threadIdx.x; i < fan; i += kThreads) s.cur[i] = 0;
    __syncthreads();
    uint64_t parent, beg, end;
    tile_range(parent_off, tile_off, n_parents, kTileKeys, parent, beg, end);
    const uint32_t fm = fan - 1;
    uint64_t x[32];
    uint32_t side = 0, kept = 0;
#pragma unroll
    for (int u = 0; u < 32; ++u) {
        uint64_t i = beg + (uint64_t)u * kThreads + threadIdx.x;
        x[u] = i < end ? ld_nc(keys + i) : kEmpty;
        if (COUNT_SIDE && i < end) {
            kept++;
            side += (x[u] == kEmpty);
        }
    }
#pragma unroll
    for (int u = 0; u < 32; ++u)
        if (x[u] != kEmpty) atomicAdd(&s.cur[digit_of(mix64(x[u]), shift, fm)], 1u);
    __syncthreads();
    const uint32_t total = scatter_plan(s, fan, child_off + parent * fan, child_cur + parent * fan);
#pragma unroll
    for (int u = 0; u < 32; ++u)
        if (x[u] != kEmpty) {
            uint32_t d = digit_of(mix64(x[u]), shift, fm);
            uint32_t pos = atomicAdd(&s.cur[d], 1u);
            stage[pos] = x[u];
            s.dig[pos] = (uint16_t)d;
        }
    __syncthreads();
    scatter_flush(s, stage, total, out);
    if (COUNT_SIDE) {
        kept = warp_sum32(kept);
        side = warp_sum32(side);
        if ((threadIdx.x & 31) == 0) {
            if (kept) atomicAdd(&ctr[C_TOTAL], (unsigned long long)kept);
            if (side) atomicAdd(&ctr[C_SIDE], (unsigned long long)side);
        }
    }
}

/* ---- count: one bucket at a time in a shared-memory table ------------------------------------- */
template <bool EMIT>
__global__ void __launch_bounds__(kThreads) k_count_buckets(const uint64_t *__restrict__ keys,
                                                            const uint64_t *__restrict__ bucket_off,
                                                            uint64_t n_buckets, Slot *__restrict__ spill,
                                                            uint64_t spill_cap,
                                                            unsigned long long *__restrict__ ctr,
                                                            uint64_t *__restrict__ out_kmers,
                                                            uint64_t *__restrict__ out_counts)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *tk = reinterpret_cast<unsigned long long *>(smem_raw);                /* keys   */
    uint32_t *tc = reinterpret_cast<uint32_t *>(smem_raw + sizeof(uint64_t) * kBucketSlots); /* extras */
    __shared__ unsigned long long row_base;
    Tally ty = {0, 0, 0, 0}; /* spilled keys account themselves through hash_insert */
    uint32_t distinct = 0, unique = 0;
    for (uint64_t b = blockIdx.x; b < n_buckets; b += gridDim.x) {
        const uint64_t beg = bucket_off[b], end = bucket_off[b + 1];
        if (beg == end) continue; /* uniform for the CTA */
        for (int i = threadIdx.x; i < kBucketSlots; i += kThreads) {
            tk[i] = kEmpty;
            tc[i] = 0;
        }
        __syncthreads();
        for (uint64_t i = beg + threadIdx.x; i < end; i += kThreads) {
            const uint64_t x = ld_nc(keys + i);
            uint32_t sl = (uint32_t)((mix64(x) * kSlotMul) >> 52); /* 12 bits, independent of the digits */
            int probes = 0;
            for (;;) {
                unsigned long long old = atomicCAS(&tk[sl], (unsigned long long)kEmpty, (unsigned long long)x);
                if (old == kEmpty) break;
                if (old == x) {
                    atomicAdd(&tc[sl], 1u);
                    break;
                }
                sl = (sl + 1) & (kBucketSlots - 1);
                if (++probes == kBucketSlots) { /* table full of other keys: count it in HBM */
                    if (!EMIT) hash_insert(spill, spill_cap, x, ty, ctr); /* EMIT re-runs: spill rows come from the table */
                    break;
                }
            }
        }
        __syncthreads();
        uint32_t mine = 0;
        for (int i = threadIdx.x; i < kBucketSlots; i += kThreads)
            if (tk[i] != kEmpty) {
                mine++;
                unique += (tc[i] == 0);
            }
        distinct += mine;
        if (EMIT) {
            uint32_t total;
            uint32_t rank = block_exscan(mine, &total);
            if (threadIdx.x == 0) row_base = atomicAdd(&ctr[C_CURSOR], (unsigned long long)total);
            __syncthreads();
            uint64_t pos = row_base + rank;
            for (int i = threadIdx.x; i < kBucketSlots; i += kThreads)
                if (tk[i] != kEmpty) {
                    out_kmers[pos] = tk[i];
                    out_counts[pos] = (uint64_t)tc[i] + 1;
                    pos++;
                }
        }
        __syncthreads();
    }
    distinct = warp_sum32(distinct);
    unique = warp_sum32(unique);
    if ((threadIdx.x & 31) == 0) {
        if (distinct) atomicAdd(&ctr[C_DISTINCT], (unsigned long long)distinct);
        if (unique) atomicAdd(&ctr[C_UNIQUE], (unsigned long long)unique);
    }
    ty.total = 0; /* totals were taken by the scatter pass */
    tally_flush(ty, ctr);
}

/* ---- multi-CTA exclusive scan (bucket offsets: up to 4 M entries) ------------------------------ */
constexpr int kScanPer = 8; /* 1024 threads x 8 = 8192 entries per CTA */
__global__ void __launch_bounds__(1024) k_scan_block(const uint64_t *__restrict__ in, uint64_t n,
                                                     uint64_t *__restrict__ out,
                                                     uint64_t *__restrict__ block_sums)
{
    __shared__ uint64_t wtot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t base = ((uint64_t)blockIdx.x * 1024 + threadIdx.x) * kScanPer;
    uint64_t v[kScanPer], sum = 0;
#pragma unroll
    for (int i = 0; i < kScanPer; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        sum += v[i];
    }
    uint64_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wtot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint64_t t = wtot[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += y;
        }
        wtot[lane] = ti - t;
        if (lane == 31) block_sums[blockIdx.x] = ti;
    }
    __syncthreads();
    uint64_t ex = wtot[wid] + inc - sum;
#pragma unroll
    for (int i = 0; i < kScanPer; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}

__global__ void __launch_bounds__(1024) k_scan_add(uint64_t *__restrict__ out, uint64_t n,
                                                   const uint64_t *__restrict__ block_off)
{
    const uint64_t add = block_off[blockIdx.x];
    const uint64_t base = ((uint64_t)blockIdx.x * 1024 + threadIdx.x) * kScanPer;
#pragma unroll
    for (int i = 0; i < kScanPer; ++i)
        if (base + i < n) out[base + i] += add;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_off[gridDim.x];
}

} /* namespace dnagpu */
