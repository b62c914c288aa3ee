// tools/probe_gather.cu — micro-probe: what does a random 16-byte gather over a 16 GB table cost on
// B200, per load flavour and L2 fetch-granularity setting? (Not part of the product; results are
// recorded in profiles/.) Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe_gather probe_gather.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

template <int V> __device__ __forceinline__ uint4 ld16(const uint4* p, u64 pol)
{
    uint4 r;
    if (V == 0) r = *p;
    else if (V == 1) r = __ldcs(p);
    else if (V == 2) r = __ldcg(p);
    else if (V == 3) r = __ldg(p);
    else if (V == 4) asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    else if (V == 5) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (V == 6) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (V == 7) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    else if (V == 8) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <int V, int U> __global__ void __launch_bounds__(256) k_gather(const uint4* tab, u64 n, u64 events, unsigned* sink)
{
    u64 pol = 0;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    unsigned acc = 0;
    for (u64 base = (u64)blockIdx.x * 256 * U; base < events; base += (u64)gridDim.x * 256 * U) {
        uint4 r[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { u64 i = base + j * 256 + threadIdx.x; u64 e = __umul64hi(mix(i * 0x9E3779B97F4A7C15ULL + 1), n); r[j] = ld16<V>(tab + e, pol); }
#pragma unroll
        for (int j = 0; j < U; ++j) acc += r[j].x ^ r[j].y ^ r[j].z ^ r[j].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
// 8-byte random reads / RED.MAX on a small (L2-sized) array
template <int MODE> __global__ void __launch_bounds__(256) k_small(u64* arr, u64 n, u64 events, unsigned* sink)
{
    u64 acc = 0;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < events; i += (u64)gridDim.x * 256) {
        u64 e = __umul64hi(mix(i * 0x9E3779B97F4A7C15ULL + 7), n);
        if (MODE == 0) acc += __ldcg(arr + e);
        else atomicMax(arr + e, i);
    }
    if (acc == 0x12345678u) *sink = (unsigned)acc;
}

template <typename F> float timeit(F f, int reps = 3)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main(int argc, char** argv)
{
    const u64 n = argc > 1 ? strtoull(argv[1], 0, 10) : 1000000000ull, events = 150000000ull;
    int gran = argc > 2 ? atoi(argv[2]) : 0;
    size_t lim = 0;
    cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity before: %zu\n", lim);
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set %d -> %s\n", gran, cudaGetErrorString(e)); }
    cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity now: %zu\n", lim);
    uint4* tab; CK(cudaMalloc(&tab, n * 16)); CK(cudaMemset(tab, 1, n * 16));
    unsigned* sink; CK(cudaMalloc(&sink, 4));
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const char* names[] = {"plain", "ldcs", "ldcg", "ldg(nc)", "noalloc+evict_first", "volatile", "L1::no_allocate", "nc+noalloc+evict_first", "relaxed.gpu"};
#define RUN(V, U, CT) { float ms = timeit([&] { k_gather<V, U><<<sm * CT, 256>>>(tab, n, events, sink); }); printf("gather %-24s U=%d ctas/sm=%d : %.3f ms  %.1f Gev/s  %.0f GB/s(32B sectors)\n", names[V], U, CT, ms, events / ms / 1e6, events * 32.0 / ms / 1e6); }
    RUN(0, 4, 8) RUN(1, 4, 8) RUN(2, 4, 8) RUN(3, 4, 8) RUN(4, 4, 8) RUN(5, 4, 8) RUN(6, 4, 8) RUN(7, 4, 8) RUN(8, 4, 8)
    RUN(2, 1, 8) RUN(2, 2, 8) RUN(2, 8, 8) RUN(2, 8, 4) RUN(2, 16, 4) RUN(2, 4, 4) RUN(2, 2, 4)
    const u64 small_n = 5000512;
    u64* arr; CK(cudaMalloc(&arr, small_n * 8 * 3)); CK(cudaMemset(arr, 0, small_n * 8 * 3));
    { float ms = timeit([&] { k_small<0><<<sm * 8, 256>>>(arr, small_n, events, sink); }); printf("random 8B read, 40MB array : %.3f ms  %.1f Gev/s\n", ms, events / ms / 1e6); }
    { float ms = timeit([&] { k_small<1><<<sm * 8, 256>>>(arr, small_n, events, sink); }); printf("random RED.MAX.64, 40MB array : %.3f ms  %.1f Gev/s\n", ms, events / ms / 1e6); }
    { float ms = timeit([&] { k_small<0><<<sm * 8, 256>>>(arr, small_n * 3, events, sink); }); printf("random 8B read, 120MB array : %.3f ms  %.1f Gev/s\n", ms, events / ms / 1e6); }
    { float ms = timeit([&] { k_small<1><<<sm * 8, 256>>>(arr, small_n * 3, events, sink); }); printf("random RED.MAX.64, 120MB array : %.3f ms  %.1f Gev/s\n", ms, events / ms / 1e6); }
    return 0;
}
