// Micro-benchmarks of the SM-side primitives the partition kernels lean on (B200, sm_100a):
// returning vs non-returning shared-memory atomics, 64-bit CAS, MATCH.ANY, non-atomic RMW.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int T = 512, ITER = 4096, BINS = 1024;

__device__ __forceinline__ uint32_t rnd(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(T) k(uint64_t *out)
{
    __shared__ uint32_t cnt[BINS * 2];
    __shared__ unsigned long long tab[4096];
    __shared__ uint16_t wcnt[16][BINS];
    for (int i = threadIdx.x; i < BINS * 2; i += T) cnt[i] = 0;
    for (int i = threadIdx.x; i < 4096; i += T) tab[i] = ~0ull;
    for (int i = threadIdx.x; i < 16 * BINS; i += T) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1, acc = 0;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 8
    for (int it = 0; it < ITER; ++it) {
        uint32_t d = rnd(s) & (BINS - 1);
        if (MODE == 0) acc += atomicAdd(&cnt[d], 1u);                 // returning ATOMS.ADD
        if (MODE == 1) atomicAdd(&cnt[d], 1u);                        // RED (result unused)
        if (MODE == 2) {                                              // ATOMS.CAS.64, mostly succeeding
            unsigned long long key = ((unsigned long long)rnd(s) << 24) | it;
            acc += (uint32_t)atomicCAS(&tab[(rnd(s)) & 4095], ~0ull, key);
        }
        if (MODE == 3) acc += __match_any_sync(0xffffffffu, d);       // MATCH.ANY
        if (MODE == 4) {                                              // warp-private rank: match + non-atomic RMW
            uint32_t peers = __match_any_sync(0xffffffffu, d);
            int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) { base = wcnt[w][d]; wcnt[w][d] = (uint16_t)(base + __popc(peers)); }
            base = __shfl_sync(0xffffffffu, base, leader);
            acc += base + __popc(peers & ((1u << lane) - 1));
        }
        if (MODE == 5) { uint32_t v = cnt[d]; cnt[d] = v + 1; acc += v; }    // plain LDS + STS random
        if (MODE == 6) { tab[d * 4 + (it & 3)] = acc; acc += (uint32_t)tab[(d * 4 + 1) & 4095]; } // 64-bit STS+LDS random
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
void run(const char *name, uint64_t *d_out, int sms)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = sms * 2;
    k<MODE><<<grid, T>>>(d_out);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) k<MODE><<<grid, T>>>(d_out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = 5.0 * grid * T * (double)ITER;
    double clk = 1.965e9;
    printf("%-44s %8.1f Gop/s  %6.3f lane-ops/clk/SM\n", name, ops / ms / 1e6, ops / (ms * 1e-3) / sms / clk);
}

int main()
{
    uint64_t *d; cudaMalloc(&d, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, 2 CTAs x 512 threads per SM, random addresses over 1024 bins\n", p.name, sms);
    run<0>("ATOMS.ADD returning (u32)", d, sms);
    run<1>("RED shared (atomicAdd, result unused)", d, sms);
    run<2>("ATOMS.CAS.64", d, sms);
    run<3>("MATCH.ANY", d, sms);
    run<4>("MATCH.ANY + leader LDS/STS.U16 + SHFL (rank)", d, sms);
    run<5>("LDS + STS u32 random (non-atomic RMW)", d, sms);
    run<6>("STS.64 + LDS.64 random", d, sms);
    return 0;
}
