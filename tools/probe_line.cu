// tools/probe_line.cu — micro-probe: the memory-system ceilings of the line sampler's access pattern on B200, as a ladder
// that adds one stream of the traversal kernel (abnn_b200/csrc/traversal.cu:k_traverse_line) at a time and none of its
// arithmetic. Each rung reports ms per 150M events (18.75M random 128-byte lines of a 16 GB table) and events/s:
//   L0  random 128-byte line reads into registers (8 lanes x LDG.128 per line)
//   L1  the kernel's staging: 32 lines per warp through cp.async into a 4 KB shared-memory stage, 32 warps per SM
//   L2  L1 + 4-byte write-back of one word per record for a fraction g of the LINES (dirty sectors -> DRAM writes)
//   L3  L2 + one random 4-byte read per record from a 20 MB array (the gate words)
//   L4  L3 + one RED.MAX.64 per line on a 40 MB array (lastVisited) + one 8-byte read per line of another 40 MB array (lastFired)
// with and without the persisting-L2 window over the two hot arrays. Not part of the product; results go to profiles/.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o probe_line probe_line.cu
// Run:   ./probe_line [records=1000000000] [neurons=5000512] [g_percent=27]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
__device__ __forceinline__ void cp_async16(u32 dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct Arrays { uint4* tab; u64 n_lines; u32* gate; u64* visited; u64* live; u64 n_neuron; u32 g_thresh; uint2* dstw; u32* fire32; u32* vis32; };

// L0: lane l reads 16 bytes of line (chunk * 32 + iteration * 4 + l / 8): 4 whole lines per warp instruction, U instructions in flight
template <int U> __global__ void __launch_bounds__(256) k_line_regs(Arrays a, u64 n_chunks, unsigned* sink)
{
    const u32 lane = threadIdx.x & 31, rec = lane & 7, sub = lane >> 3;
    const u64 warps = (u64)gridDim.x * 8, w = (u64)blockIdx.x * 8 + (threadIdx.x >> 5);
    unsigned acc = 0;
    for (u64 c = w; c < n_chunks; c += warps) {
#pragma unroll
        for (int k0 = 0; k0 < 8; k0 += U) {
            uint4 r[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const u64 line = __umul64hi(mix((c * 32 + (k0 + k) * 4 + sub) * 0x9E3779B97F4A7C15ULL + 1), a.n_lines);
                r[k] = __ldcg(a.tab + line * 8 + rec);
            }
#pragma unroll
            for (int k = 0; k < U; ++k) acc += r[k].x ^ r[k].w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// L0b: random blocks of BL consecutive 128-byte lines (sample_block = 8*BL records per draw): DRAM page locality vs block size
template <int BL> __global__ void __launch_bounds__(256) k_block_regs(Arrays a, u64 n_chunks, unsigned* sink)
{
    const u32 lane = threadIdx.x & 31, rec = lane & 7, sub = lane >> 3;
    const u64 warps = (u64)gridDim.x * 8, w = (u64)blockIdx.x * 8 + (threadIdx.x >> 5);
    unsigned acc = 0;
    for (u64 c = w; c < n_chunks; c += warps) {
#pragma unroll
        for (int k0 = 0; k0 < 8; k0 += 4) {
            uint4 r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u64 idx = c * 32 + (k0 + k) * 4 + sub;                 // line ordinal of the pass
                const u64 blk = __umul64hi(mix((idx / BL) * 0x9E3779B97F4A7C15ULL + 1), a.n_lines / BL);
                r[k] = __ldcg(a.tab + (blk * BL + idx % BL) * 8 + rec);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) acc += r[k].x ^ r[k].w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// L1..L4: the kernel's structure — 8 warps per CTA, one 4 KB stage per warp, 8 cp.async per lane per chunk
// LEVEL 9 / 10: as 8, with the write-back on a fraction g of the RECORDS (every line gets some: the interleaved table) —
//   9 into the record's own 16 bytes (array-of-records line: up to 4 dirty sectors per line),
//  10 into a 32-byte weight sector of the line (records of a line stored field by field: one dirty sector per line).
// BL = 2: lines drawn in blocks of two (sample_block 16).
template <int LEVEL, int BL = 1> __global__ void __launch_bounds__(256, 4) k_line_staged(Arrays a, u64 n_chunks, unsigned* sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, rec = lane & 7, sub = lane >> 3;
    unsigned char* stage = smem + warp * 4096;
    const u32 mine = (u32)__cvta_generic_to_shared(stage + lane * 16);
    const u64 warps = (u64)gridDim.x * 8, w = (u64)blockIdx.x * 8 + warp;
    unsigned acc = 0;
    for (u64 c = w; c < n_chunks; c += warps) {
        const u64 my_line = BL == 1 ? __umul64hi(mix((c * 32 + lane) * 0x9E3779B97F4A7C15ULL + 1), a.n_lines)    // lane L draws line L
                                    : __umul64hi(mix((c * 16 + lane / 2) * 0x9E3779B97F4A7C15ULL + 1), a.n_lines / 2) * 2 + (lane & 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const u64 line = __shfl_sync(0xffffffffu, my_line, k * 4 + sub);
            cp_async16(mine + k * 512, a.tab + line * 8 + rec);
        }
        cp_async_commit();
        cp_async_wait0();
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 r = *reinterpret_cast<const uint4*>(stage + lane * 16 + k * 512);
            const u64 line = __shfl_sync(0xffffffffu, my_line, k * 4 + sub);
            if (LEVEL >= 3) acc += __ldcg(a.gate + r.x);                                      // one random 4-byte read per record
            if (LEVEL == 5 && rec == 7) {                                                     // packed {vis32, fire32}: one sector per line
                const u32 dst = (u32)__umul64hi(mix(line), a.n_neuron);
                acc += __ldcg(reinterpret_cast<const u32*>(a.dstw + dst) + 1);
                atomicMax(reinterpret_cast<u32*>(a.dstw + dst), (u32)(c * 256 + k * 32 + lane));
            }
            if (LEVEL == 6 && rec == 7) {                                                     // two 20 MB arrays (fire32, vis32)
                const u32 dst = (u32)__umul64hi(mix(line), a.n_neuron);
                acc += __ldcg(a.fire32 + dst);
                atomicMax(a.vis32 + dst, (u32)(c * 256 + k * 32 + lane));
            }
            if (LEVEL >= 8) {                                                                 // interleaved order: the 8 records of a line
                const u32 dst = ((u32)__umul64hi(mix(line), a.n_neuron - 8) & ~7u) + rec;     // target 8 ADJACENT neurons (one sector)
                acc += __ldcg(a.fire32 + dst);
                atomicMax(a.vis32 + dst, (u32)(c * 256 + k * 32 + lane));
            }
            if (LEVEL == 7 && rec == 7) {                                                     // packed, read only (no RED)
                const u32 dst = (u32)__umul64hi(mix(line), a.n_neuron);
                acc += __ldcg(reinterpret_cast<const u32*>(a.dstw + dst) + 1);
            }
            if (LEVEL == 4 && rec == 7) {                                                     // once per line
                const u32 dst = (u32)__umul64hi(mix(line), a.n_neuron);
                acc += (unsigned)__ldcg(a.live + dst);
                atomicMax(a.visited + dst, c * 256 + k * 32 + lane);
            }
            if (LEVEL >= 9) {
                if ((u32)mix((line * 8 + rec) ^ (c * 0x5bd1e995u)) < a.g_thresh) {             // a fraction g of the records
                    if (LEVEL == 9) __stcg(reinterpret_cast<u32*>(a.tab + line * 8 + rec) + 2, r.z + 1u);
                    else            __stcg(reinterpret_cast<u32*>(a.tab + line * 8) + 16 + rec, r.x);   // (a valid gate index: bytes 64..95 hold records 4 and 5)
                }
            } else
            if (LEVEL >= 2 && (u32)mix(line ^ 0x5bd1e995u) < a.g_thresh)                     // a fraction g of the lines is written back
                __stcg(reinterpret_cast<u32*>(a.tab + line * 8 + rec) + 2, r.z + 1u);
            acc += r.y;
        }
        __syncwarp();
    }
    if (acc == 0x12345678u) *sink = acc;
}

// I0..I3: the iid sampler's streams (k_traverse_line32<.,.,1>): every EVENT gathers its own random 16-byte record with cp.async
// (256 per warp and chunk); LEVEL 1 + one gate read per record, 2 + fire32 read and vis32 RED per record, 3 + 4-byte
// write-back on a fraction g of the records
template <int LEVEL> __global__ void __launch_bounds__(256, 4) k_iid_staged(Arrays a, u64 n_chunks, unsigned* sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* stage = smem + warp * 4096;
    const u32 mine = (u32)__cvta_generic_to_shared(stage + lane * 16);
    const u64 warps = (u64)gridDim.x * 8, w = (u64)blockIdx.x * 8 + warp, n_rec = a.n_lines * 8;
    unsigned acc = 0;
    for (u64 c = w; c < n_chunks; c += warps) {
        u64 recs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            recs[k] = __umul64hi(mix((c * 256 + k * 32 + lane) * 0x9E3779B97F4A7C15ULL + 1), n_rec);
            cp_async16(mine + k * 512, a.tab + recs[k]);
        }
        cp_async_commit();
        cp_async_wait0();
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 r = *reinterpret_cast<const uint4*>(stage + lane * 16 + k * 512);
            if (LEVEL >= 1) acc += __ldcg(a.gate + r.x);
            if (LEVEL >= 2) {
                const u32 dst = (u32)__umul64hi(mix(recs[k]), a.n_neuron);
                acc += __ldcg(a.fire32 + dst);
                atomicMax(a.vis32 + dst, (u32)(c * 256 + k * 32 + lane));
            }
            if (LEVEL >= 3 && (u32)mix(recs[k] ^ (c * 0x5bd1e995u)) < a.g_thresh)
                __stcg(reinterpret_cast<u32*>(a.tab + recs[k]) + 2, r.z + 1u);
            acc += r.y;
        }
        __syncwarp();
    }
    if (acc == 0x12345678u) *sink = acc;
}

// table records: x = uniformly random source neuron (the gate-word index), y = z = w = filler
__global__ void k_init_table(uint4* tab, u64 n, u64 n_neuron)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        tab[i] = make_uint4((u32)(mix(i * 0xD1B54A32D192ED03ULL + 3) % n_neuron), (u32)i, 0x3e4ccccdu, 0u);
}

template <typename F> float timeit(F f, int reps = 5)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main(int argc, char** argv)
{
    const u64 n = argc > 1 ? strtoull(argv[1], 0, 10) : 1000000000ull;
    const u64 neurons = argc > 2 ? strtoull(argv[2], 0, 10) : 5000512ull;
    const double g = (argc > 3 ? atof(argv[3]) : 27.0) / 100.0;
    const u64 events = 150000000ull, n_chunks = events / 256;
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    Arrays a{};
    a.n_lines = n / 8; a.n_neuron = neurons; a.g_thresh = (u32)(g * 4294967295.0);
    CK(cudaMalloc(&a.tab, n * 16));
    k_init_table<<<sm * 16, 256>>>(a.tab, n, neurons);
    CK(cudaDeviceSynchronize());
    // hot arrays in one allocation, hottest first, like the product: [gate 4 B | visited 8 B | live 8 B] per neuron
    const u64 npad = (neurons + 31) & ~31ull;
    unsigned char* hot; CK(cudaMalloc(&hot, npad * 20)); CK(cudaMemset(hot, 0, npad * 20));
    a.gate = reinterpret_cast<u32*>(hot); a.visited = reinterpret_cast<u64*>(hot + npad * 4); a.live = a.visited + npad;
    // second layout: [gate 4 B | dstw {vis32, fire32} 8 B] per neuron = 60 MB, and [gate | fire32 | vis32]
    unsigned char* hot2; CK(cudaMalloc(&hot2, npad * 12)); CK(cudaMemset(hot2, 0, npad * 12));
    u32* gate2 = reinterpret_cast<u32*>(hot2 + npad * 8);
    a.dstw = reinterpret_cast<uint2*>(hot2); a.fire32 = reinterpret_cast<u32*>(hot2); a.vis32 = a.fire32 + npad;
    unsigned* sink; CK(cudaMalloc(&sink, 4));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    CK(cudaFuncSetAttribute(k_line_staged<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608));
    CK(cudaFuncSetAttribute(k_line_staged<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608));
    CK(cudaFuncSetAttribute(k_line_staged<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608));
    CK(cudaFuncSetAttribute(k_line_staged<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608));
    CK(cudaFuncSetAttribute(k_line_staged<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608));
    CK(cudaFuncSetAttribute(k_line_staged<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608));
    CK(cudaFuncSetAttribute(k_line_staged<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608));
    CK(cudaFuncSetAttribute(k_line_staged<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608));
    CK((cudaFuncSetAttribute(k_line_staged<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608)));
    CK((cudaFuncSetAttribute(k_line_staged<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608)));
    CK((cudaFuncSetAttribute(k_line_staged<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608)));
    CK((cudaFuncSetAttribute(k_line_staged<9, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608)));
    CK((cudaFuncSetAttribute(k_line_staged<10, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608)));
    const bool quick = argc > 4;                              // only the write-pattern rungs (L8 / L9 / L10)
    auto report = [&](const char* name, float ms, double bytes_per_event) {
        printf("%-64s %.3f ms  %6.1f Gev/s  %5.2f TB/s algorithmic\n", name, ms, events / ms / 1e6, events * bytes_per_event / ms / 1e9);
    };
    printf("table %.1f GB (%llu records), %llu neurons, g = %.2f, %d SMs\n", n * 16 / 1e9, n, neurons, g, sm);
    if (!quick) {
    report("L0 line reads into registers, 2 x 4 lines in flight per warp", timeit([&] { k_line_regs<2><<<sm * 8, 256, 0, st>>>(a, n_chunks, sink); }), 16);
    report("L0 line reads into registers, 4 x 4 lines in flight per warp", timeit([&] { k_line_regs<4><<<sm * 8, 256, 0, st>>>(a, n_chunks, sink); }), 16);
    report("L0 line reads into registers, 8 x 4 lines in flight per warp", timeit([&] { k_line_regs<8><<<sm * 8, 256, 0, st>>>(a, n_chunks, sink); }), 16);
    report("L0b blocks of 2 lines (256 B) per draw", timeit([&] { k_block_regs<2><<<sm * 8, 256, 0, st>>>(a, n_chunks, sink); }), 16);
    report("L0b blocks of 4 lines (512 B) per draw", timeit([&] { k_block_regs<4><<<sm * 8, 256, 0, st>>>(a, n_chunks, sink); }), 16);
    report("L0b blocks of 8 lines (1 KB) per draw", timeit([&] { k_block_regs<8><<<sm * 8, 256, 0, st>>>(a, n_chunks, sink); }), 16);
    report("L0b blocks of 32 lines (4 KB) per draw", timeit([&] { k_block_regs<32><<<sm * 8, 256, 0, st>>>(a, n_chunks, sink); }), 16);
    }
    for (int window = 0; window < 2 && !quick; ++window) {
        if (window) {                       // persisting window over [gate | visited] = 12 B per neuron, like the product
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, 0);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, 0);
            size_t hotb = npad * 12, want = hotb < (size_t)max_persist ? hotb : (size_t)max_persist;
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = hot;
            attr.accessPolicyWindow.num_bytes = hotb < (size_t)max_window ? hotb : (size_t)max_window;
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
            printf("-- persisting L2 window over gate words + lastVisited (%.0f MB of %.0f MB set aside)\n", hotb / 1e6, want / 1e6);
        }
        const size_t sh = 8 * 4608;
        report("L1 staging only (cp.async, 4 KB per warp, 32 warps per SM)", timeit([&] { k_line_staged<1><<<sm * 4, 256, sh, st>>>(a, n_chunks, sink); }), 16);
        report("L2 + 4-byte write-back on a fraction g of the lines", timeit([&] { k_line_staged<2><<<sm * 4, 256, sh, st>>>(a, n_chunks, sink); }), 16 + 16 * g);
        report("L3 + one random 4-byte gate read per record (20 MB array)", timeit([&] { k_line_staged<3><<<sm * 4, 256, sh, st>>>(a, n_chunks, sink); }), 16 + 16 * g);
        report("L4 + RED.MAX.64 and 8-byte read per line (two 40 MB arrays)", timeit([&] { k_line_staged<4><<<sm * 4, 256, sh, st>>>(a, n_chunks, sink); }), 16 + 16 * g);
    }
    // ---- packed 32-bit destination words: [gate | {vis32, fire32}] = 60 MB
    Arrays b = a; b.gate = gate2;
    for (int window = 0; window < 5; ++window) {
        cudaStreamAttrValue attr{};
        if (window >= 3) {
            int max_window = 0;
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, 0);
            const size_t hotb = window == 3 ? npad * 8 : npad * 12, want = window == 3 ? npad * 8 : npad * 10;
            CK(cudaCtxResetPersistingL2Cache());
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
            attr.accessPolicyWindow.base_ptr = hot2;
            attr.accessPolicyWindow.num_bytes = hotb < (size_t)max_window ? hotb : (size_t)max_window;
            attr.accessPolicyWindow.hitRatio = window == 3 ? 1.0f : 0.83f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            printf("-- [fire32 | vis32 | gate]: %s\n", window == 3 ? "window over fire32 + vis32 only (40 MB set aside)" : "window over all three, 50 MB set aside, hit ratio 0.83");
        } else if (window) {
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, 0);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, 0);
            size_t hotb = npad * 12, want = hotb < (size_t)max_persist ? hotb : (size_t)max_persist;
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
            attr.accessPolicyWindow.base_ptr = hot2;
            attr.accessPolicyWindow.num_bytes = hotb < (size_t)max_window ? hotb : (size_t)max_window;
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = window == 1 ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
            printf("-- packed layout, persisting window over gate + dstw (60 MB), missProp %s\n", window == 1 ? "streaming" : "normal");
        } else {
            CK(cudaCtxResetPersistingL2Cache());
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
            attr.accessPolicyWindow.num_bytes = 0;
            printf("-- packed layout, no persisting window\n");
        }
        CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
        const size_t sh = 8 * 4608;
        if (quick) {
            if (window != 1) continue;
            CK((cudaFuncSetAttribute(k_iid_staged<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608)));
            CK((cudaFuncSetAttribute(k_iid_staged<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608)));
            CK((cudaFuncSetAttribute(k_iid_staged<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608)));
            CK((cudaFuncSetAttribute(k_iid_staged<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608)));
            report("I0 iid: one random 16-byte record per event (cp.async, 256 per warp)", timeit([&] { k_iid_staged<0><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16);
            report("I1 + gate read per record", timeit([&] { k_iid_staged<1><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16);
            report("I2 + fire32 read and vis32 RED per record", timeit([&] { k_iid_staged<2><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16);
            report("I3 + 4-byte write-back on a fraction g of the records", timeit([&] { k_iid_staged<3><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
            for (int rep = 0; rep < 2; ++rep) {
            report("L8  write-back: all records of a fraction g of the lines", timeit([&] { k_line_staged<8><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
            report("L9  write-back: a fraction g of the records, in place", timeit([&] { k_line_staged<9><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
            report("L10 write-back: a fraction g of the records, weight sector", timeit([&] { k_line_staged<10><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
            report("L8  blocks of 2 lines", timeit([&] { k_line_staged<8, 2><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
            report("L9  blocks of 2 lines", timeit([&] { k_line_staged<9, 2><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
            report("L10 blocks of 2 lines", timeit([&] { k_line_staged<10, 2><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
            }
            continue;
        }
        report("L3 gate read per record (20 MB array)", timeit([&] { k_line_staged<3><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
        report("L7 + 4-byte read per line of packed dstw (40 MB)", timeit([&] { k_line_staged<7><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
        report("L5 + RED.MAX.32 on the same 8-byte word", timeit([&] { k_line_staged<5><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
        report("L6 fire32 / vis32 as two 20 MB arrays (read + RED.MAX.32)", timeit([&] { k_line_staged<6><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
        report("L8 same, 8 adjacent neurons per line (read + RED per record)", timeit([&] { k_line_staged<8><<<sm * 4, 256, sh, st>>>(b, n_chunks, sink); }), 16 + 16 * g);
    }
    return 0;
}
