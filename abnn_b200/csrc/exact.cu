// abnn_b200/csrc/exact.cu — EXACT execution: parallel, and bit-identical to the SERIAL event order.
//
// With the SNAPSHOT src view and the once-per-pass r-bar step, two events of a pass interact only
// through their destination neuron: lastFired[dst] (refractory gate, inter-spike interval, fire
// timestamp) and — the same edge implies the same dst — the synapse weight. lastVisited is an
// order-free max. So the pass splits into
//   phase 1  (parallel over events)      : sample, gather, lastVisited RED, pre-spike window test against
//                                          the snapshot; survivors ("candidates") are emitted as
//                                          key = (dst << 32 | event index), value = edge index;
//   sort     (cub::DeviceRadixSort)      : groups candidates by destination, event order inside a group;
//   phase 3  (parallel over destinations): one thread walks one destination's candidates in event order
//                                          with lastFired[dst] in a register — exactly the serial loop
//                                          restricted to that neuron.
// The global spike budget (brain.metal:85-98) is an ordered prefix over ALL events: the k-th fire in event order closes
// the gate for every later event of every destination, which couples all the per-destination chains of phase 3. It is
// not supported here (max_spikes_per_pass must be 0): SERIAL execution carries the budgeted metal-parity profile
// bit-exactly, PARALLEL execution honours the budget with a saturating counter (tests/test_gpu_line32.py).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "kernels.h"

namespace abnn {

namespace {

__device__ __forceinline__ bool sample_edge(const KParams& kp, u64 event_base, u64 i, u64* edge)
{
    if (kp.sampler == ABNN_SAMPLER_SWEEP) { *edge = i; return i < kp.n_local; }
    if (!kp.n_local) return false;
    const u64 lane = i & (kp.sample_block - 1);
    const u64 eid0 = event_base + i - lane;
    const Philox4 q = philox4x32_10((u32)eid0, (u32)(eid0 >> 32), kp.rank, STREAM_EVENT, kp.seed_lo, kp.seed_hi);
    const u64 e = kp.sample_block == 1 ? mulhi64(((u64)q.x << 32) | q.y, kp.n_local)
                                       : (mulhi64(((u64)q.x << 32) | q.y, kp.n_blocks) << kp.log_block) + lane;
    *edge = e;
    return e < kp.n_local;
}

__global__ void __launch_bounds__(256) k_exact_phase1(const __grid_constant__ KParams kp, const DevPtrs d, u64* keys, u64* vals,
                                                      u32* counter)
{
    const u64 clock = d.sc->clock, event_base = d.sc->event_base;
    const unsigned lane = threadIdx.x & 31;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 rounds = (kp.count + stride - 1) / stride;
    for (u64 r = 0; r < rounds; ++r) {                       // whole warps iterate together (ballot below)
        const u64 i = r * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        bool cand = false;
        u64 edge = 0;
        u32 dst = 0;
        if (i < kp.count && sample_edge(kp, event_base, i, &edge)) {
            const uint4 s = __ldcs(reinterpret_cast<const uint4*>(d.syn + edge));
            if (s.x != DEAD_SRC) {                                   // a dead record waits for the next rebuild: no event
                const u64 now = kp.clock_mode == ABNN_CLOCK_PER_PASS ? clock : clock + i * kp.world + kp.rank;
                if (kp.track_visits) atomicMax(d.visited + s.y, now);
                const u64 lp = __ldcg(d.view + s.x);
                cand = now - lp <= kp.window_pre;
                dst = s.y;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, cand);
        if (!m) continue;
        u32 base = 0;
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(counter, (u32)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (cand) {
            const u32 slot = base + __popc(m & ((1u << lane) - 1u));
            keys[slot] = ((u64)dst << 32) | (u32)i;
            vals[slot] = edge;
        }
    }
}

__global__ void __launch_bounds__(256) k_exact_phase3(const __grid_constant__ KParams kp, const DevPtrs d, const u64* __restrict__ keys,
                                                      const u64* __restrict__ vals, const u32* __restrict__ n_ptr)
{
    __shared__ u32 s_cnt[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u32 n = *n_ptr;
    const u64 clock = d.sc->clock, event_base = d.sc->event_base, tick_base = d.sc->tick_base;
    const float R = d.sc->reward, rbar = d.sc->rbar;
    u32 gated = 0, fired_n = 0;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (u64)gridDim.x * blockDim.x) {
        const u32 dst = (u32)(keys[j] >> 32);
        if (j > 0 && (u32)(keys[j - 1] >> 32) == dst) continue;           // not the head of a destination's chain
        u64 ld = d.live[dst];
        const u64 ld0 = ld;
        for (u64 t = j; t < n; ++t) {
            const u64 key = keys[t];
            if ((u32)(key >> 32) != dst) break;
            const u64 i = (u32)key;
            const u64 now = kp.clock_mode == ABNN_CLOCK_PER_PASS ? clock : clock + i * kp.world + kp.rank;
            if (now - ld <= kp.refractory) continue;                       // brain.metal:79-83
            const u64 edge = vals[t];
            const float w = d.syn[edge].w;
            const u64 eid = event_base + i;
            const u32 group = kp.sampler == ABNN_SAMPLER_PHILOX ? kp.sample_block : 1u;      // events per Philox call
            const u32 glane = (u32)i & (group - 1u);
            Philox4 r{0, 0, 0, 0};
            if (kp.release_rng == ABNN_RNG_PHILOX || kp.p_new > 0.f) {
                const u64 eid0 = eid - glane;
                r = philox4x32_10((u32)eid0, (u32)(eid0 >> 32), kp.rank, STREAM_EVENT, kp.seed_lo, kp.seed_hi);
            }
            const float u = kp.release_rng == ABNN_RNG_XORSHIFT ? rand01_xorshift((u32)i ^ (u32)now) : u01_24(release_word(r.z, group, glane));
            const float p = clampf(w * w * kp.base_scale, 0.f, 1.f);       // brain.metal:91
            const bool fired = p > u;                                      // brain.metal:92
            float dW = fired ? (kp.a_ltp * (1.f - w)) : (-kp.a_ltd * w);   // brain.metal:101-102
            dW += kp.eta_reward * (R - rbar) * (fired ? 1.0f : 0.0f);      // brain.metal:107
            const float isi = (float)(now - ld);                           // brain.metal:116
            const float est = isi > 0.f ? kp.home_tick_hz / isi : 0.f;     // brain.metal:117
            dW += kp.eta_home * (kp.target_rate_hz - est) * w;             // brain.metal:118
            const float w_new = clampf(w + dW, kp.w_min, kp.w_max);        // brain.metal:121
            d.syn[edge].w = w_new;                                         // brain.metal:122
            stage_prune(kp, d, edge, w_new);
            ++gated;
            if (fired) {
                if (ld < now) ld = now;                                    // brain.metal:125-126
                ++fired_n;
                if (kp.p_new > 0.f && (float)trial_word(r.w, group, glane) * (1.0f / 4294967296.0f) < kp.p_new) {   // README.md:125
                    const Philox4 g = philox4x32_10((u32)eid, (u32)(eid >> 32), kp.rank, STREAM_GROW, kp.seed_lo, kp.seed_hi);
                    const u32 nd = (u32)(kp.n_input + mulhi64(((u64)g.x << 32) | g.y, kp.n_neuron - kp.n_input));
                    const u32 slot = atomicAdd(&d.sc->grow_count, 1u);
                    if (slot < kp.grow_cap) d.grow[slot] = GrowCand{tick_base + i * kp.world + kp.rank, d.syn[edge].src, nd};
                    else atomicAdd(&d.sc->grow_overflow, 1u);
                    atomicAdd(&d.sc->grown_pass, 1ull);
                }
            }
        }
        if (ld != ld0) d.live[dst] = ld;
    }
    gated = __reduce_add_sync(0xffffffffu, gated);
    fired_n = __reduce_add_sync(0xffffffffu, fired_n);
    if ((threadIdx.x & 31) == 0) { if (gated) atomicAdd(&s_cnt[0], gated); if (fired_n) atomicAdd(&s_cnt[1], fired_n); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(&d.sc->gated, (u64)s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&d.sc->fired, (u64)s_cnt[1]);
        if (blockIdx.x == 0) d.sc->cands = n;
    }
}

}  // namespace

size_t exact_sort_temp_bytes(u64 cap)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const u64*)nullptr, (u64*)nullptr, (const u64*)nullptr, (u64*)nullptr,
                                    (size_t)cap, 0, 64);
    return bytes;
}

// keys/vals: 2 x cap each ([0,cap) input, [cap,2cap) sorted output).
// k_exact_pad: keys[n .. n_slots) = a key behind every neuron's, n = the candidate count phase 1 left in *counter
__global__ void __launch_bounds__(256) k_exact_pad(u64* keys, const u32* __restrict__ counter, u64 n_slots, u64 pad_key)
{
    for (u64 i = *counter + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) keys[i] = pad_key;
}
cudaError_t launch_exact_phase1(const KParams& kp, const DevPtrs& d, u64* keys, u64* vals, u32* counter, u64 n_slots, int dst_bits,
                                int sm_count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(u32), st);
    if (e != cudaSuccess || !kp.count) return e;
    u64 blocks = (kp.count + 255) / 256;
    const u64 cap = (u64)sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_exact_phase1<<<(unsigned)blocks, 256, 0, st>>>(kp, d, keys, vals, counter);
    k_exact_pad<<<(unsigned)blocks, 256, 0, st>>>(keys, counter, n_slots, ((1ull << dst_bits) << 32) | 0xFFFFFFFFull);
    return cudaGetLastError();
}
cudaError_t launch_exact_sort(u64* keys, u64* vals, u64 cap, u32 n, int key_bits, void* tmp, size_t tmp_bytes, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys + cap, vals, vals + cap, (size_t)n, 0, key_bits, st);
}
cudaError_t launch_exact_phase3(const KParams& kp, const DevPtrs& d, const u64* keys_sorted, const u64* vals_sorted,
                                const u32* counter, u32 n_host, int sm_count, cudaStream_t st)
{
    u64 blocks = ((u64)n_host + 255) / 256;
    if (blocks == 0) blocks = 1;
    const u64 cap = (u64)sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_exact_phase3<<<(unsigned)blocks, 256, 0, st>>>(kp, d, keys_sorted, vals_sorted, counter);
    return cudaGetLastError();
}

}  // namespace abnn
