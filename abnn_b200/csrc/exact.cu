// abnn_b200/csrc/exact.cu — EXACT execution: parallel, and bit-identical to the SERIAL event order.
//
// With the SNAPSHOT src view and the once-per-pass r-bar step, two events of a pass interact only
// through their destination neuron: lastFired[dst] (refractory gate, inter-spike interval, fire
// timestamp) and — the same edge implies the same dst — the synapse weight. lastVisited is an
// order-free max. So the pass splits into
//   phase 1  (parallel over events)      : sample, gather, lastVisited RED, pre-spike window test against
//                                          the snapshot ("candidates", counted), refractory test against
//                                          lastFired[dst] AS IT IS AT THE START OF THE PASS; survivors are
//                                          emitted as (dst << 32 | event index) and counted per destination;
//   group    (scan + scatter)            : exclusive scan of the per-destination counts, then every survivor's
//                                          event index goes into its destination's bucket, at the arrival number
//                                          the counting atomic of phase 1 returned (a counting sort by
//                                          destination: one pass over the survivors instead of a 7-pass radix
//                                          sort of a buffer padded to the events of the pass);
//   phase 3  (parallel over destinations): one thread per destination sorts its bucket by event index (a few
//                                          dozen entries) and walks it with lastFired[dst] in a register —
//                                          exactly the serial loop restricted to that neuron. The synapse of an
//                                          event is re-derived from its index (the Philox call it needs for the
//                                          release draw anyway).
// Dropping the events that are refractory against the pass-start lastFired is exact: lastFired[dst] only grows during a
// pass and never beyond the tick of an event already processed, so an event inside the refractory period of the
// pass-start value is inside the refractory period of whatever value the serial order would show it (a timestamp in the
// event's future wraps the unsigned difference and is kept, as the serial loop keeps it). In the benchmark regime that
// leaves 27 % of the events for the sort and the chains instead of 95 %.
// The global spike budget (brain.metal:85-98) is an ordered prefix over ALL events: the k-th fire in event order closes
// the gate for every later event of every destination, which couples all the per-destination chains of phase 3. It is
// not supported here (max_spikes_per_pass must be 0): SERIAL execution carries the budgeted metal-parity profile
// bit-exactly, PARALLEL execution honours the budget with a saturating counter (tests/test_gpu_line32.py).
#include <cub/device/device_scan.cuh>

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

// list[0 .. counter[0]) = (dst << 32 | event) of the events still open against the pass-start lastFired, in no particular
// order (one slot reservation per CTA and round, not per warp: 600k atomics on one address instead of 4.7M);
// cnt[dst - lo] = how many of them each destination got, slot[j] = the arrival number of list[j] at its destination (the
// value the counting atomic returned: the scatter needs no second round of atomics); counter[1] = events that passed the
// pre-spike window (the pass's candidate count)
__global__ void __launch_bounds__(256) k_exact_phase1(const __grid_constant__ KParams kp, const DevPtrs d, u64* list, u32* slot, u32* cnt,
                                                      u32 lo, u32* counter)
{
    __shared__ u32 s_warp[8], s_base;
    const u64 clock = d.sc->clock, event_base = d.sc->event_base;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 rounds = (kp.count + stride - 1) / stride;
    u32 n_cand = 0;
    for (u64 r = 0; r < rounds; ++r) {                       // the whole CTA iterates together (barriers below)
        const u64 i = r * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        bool open = false;
        u64 edge = 0;
        u32 dst = 0;
        if (i < kp.count && sample_edge(kp, event_base, i, &edge)) {
            const uint4 s = __ldcs(reinterpret_cast<const uint4*>(d.syn + edge));      // brain.metal:70
            if (s.x != DEAD_SRC) {                                   // a dead record waits for the next rebuild: no event
                const u64 now = kp.clock_mode == ABNN_CLOCK_PER_PASS ? clock : clock + i * kp.world + kp.rank;
                const u64 lp = __ldcg(d.view + s.x);                                   // brain.metal:73
                const u64 ld0 = __ldcg(d.live + s.y);                                  // pass-start value: phase 3 is the only writer
                if (kp.track_visits) atomicMax(d.visited + s.y, now);                  // README.md:84
                const bool cand = now - lp <= kp.window_pre;                           // brain.metal:74
                n_cand += cand;
                open = cand && !(now - ld0 <= kp.refractory);                          // brain.metal:79-83 against the pass-start value
                dst = s.y;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, open);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 total = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const u32 c = s_warp[w]; s_warp[w] = total; total += c; }
            s_base = total ? atomicAdd(counter, total) : 0u;
        }
        __syncthreads();
        if (open) {
            const u32 at = s_base + s_warp[warp] + __popc(m & ((1u << lane) - 1u));
            list[at] = ((u64)dst << 32) | (u32)i;
            slot[at] = atomicAdd(cnt + (dst - lo), 1u);
        }
        __syncthreads();                                     // s_warp / s_base are rewritten in the next round
    }
    n_cand = __reduce_add_sync(0xffffffffu, n_cand);
    if (lane == 0 && n_cand) atomicAdd(counter + 1, n_cand);
}

// start[] = exclusive scan of cnt[]: bucket n is bucket[start[n] .. start[n] + cnt[n]); every open event goes to the slot of
// its arrival number
__global__ void __launch_bounds__(256) k_exact_scatter(const u64* __restrict__ list, const u32* __restrict__ slot,
                                                       const u32* __restrict__ counter, const u32* __restrict__ start, u32 lo, u32* bucket)
{
    const u32 n = *counter;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (u64)gridDim.x * blockDim.x) {
        const u64 it = list[j];
        bucket[__ldg(start + ((u32)(it >> 32) - lo)) + slot[j]] = (u32)it;
    }
}

// Event order inside a bucket. Buckets hold a few dozen entries in the regimes the kernel is built for (insertion sort);
// a tiny network under a long pass can put thousands of events on one destination, so larger buckets take a heap sort
// (in place, O(k log k) whatever the arrival order).
__device__ __forceinline__ void sift_down(u32* a, u32 root, u32 n)
{
    const u32 v = a[root];
    for (;;) {
        u32 child = 2 * root + 1;
        if (child >= n) break;
        if (child + 1 < n && a[child + 1] > a[child]) ++child;
        if (a[child] <= v) break;
        a[root] = a[child];
        root = child;
    }
    a[root] = v;
}
__device__ void sort_bucket(u32* bk, u32 k)
{
    if (k <= 32) {
        for (u32 a = 1; a < k; ++a) {
            const u32 v = bk[a];
            u32 b = a;
            while (b > 0 && bk[b - 1] > v) { bk[b] = bk[b - 1]; --b; }
            bk[b] = v;
        }
        return;
    }
    for (u32 start = k / 2; start-- > 0;) sift_down(bk, start, k);
    for (u32 end = k - 1; end > 0; --end) {
        const u32 top = bk[0]; bk[0] = bk[end]; bk[end] = top;
        sift_down(bk, 0, end);
    }
}

__global__ void __launch_bounds__(256) k_exact_phase3(const __grid_constant__ KParams kp, const DevPtrs d, u32* bucket,
                                                      const u32* __restrict__ cnt, const u32* __restrict__ cursor, u32 lo, u32 span,
                                                      const u32* __restrict__ counter)
{
    __shared__ u32 s_cnt[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u64 clock = d.sc->clock, event_base = d.sc->event_base, tick_base = d.sc->tick_base;
    const float R = d.sc->reward, rbar = d.sc->rbar;
    u32 gated = 0, fired_n = 0;
    for (u64 nn = (u64)blockIdx.x * blockDim.x + threadIdx.x; nn < span; nn += (u64)gridDim.x * blockDim.x) {
        const u32 k = cnt[nn];
        if (!k) continue;
        const u32 dst = lo + (u32)nn;
        u32* bk = bucket + cursor[nn];
        sort_bucket(bk, k);
        u64 ld = d.live[dst];
        const u64 ld0 = ld;
        for (u32 t = 0; t < k; ++t) {
            const u64 i = bk[t];
            const u64 now = kp.clock_mode == ABNN_CLOCK_PER_PASS ? clock : clock + i * kp.world + kp.rank;
            if (now - ld <= kp.refractory) continue;                       // brain.metal:79-83
            const u64 eid = event_base + i;
            const u32 group = kp.sampler == ABNN_SAMPLER_PHILOX ? kp.sample_block : 1u;      // events per Philox call
            const u32 glane = (u32)i & (group - 1u);
            Philox4 r{0, 0, 0, 0};
            if (kp.sampler == ABNN_SAMPLER_PHILOX || kp.release_rng == ABNN_RNG_PHILOX || kp.p_new > 0.f) {
                const u64 eid0 = eid - glane;
                r = philox4x32_10((u32)eid0, (u32)(eid0 >> 32), kp.rank, STREAM_EVENT, kp.seed_lo, kp.seed_hi);
            }
            // the event's synapse again (sample_edge; phase 1 has checked that it exists)
            const u64 edge = kp.sampler == ABNN_SAMPLER_SWEEP ? i
                           : kp.sample_block == 1 ? mulhi64(((u64)r.x << 32) | r.y, kp.n_local)
                                                  : (mulhi64(((u64)r.x << 32) | r.y, kp.n_blocks) << kp.log_block) + glane;
            const float w = d.syn[edge].w;
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
        if (blockIdx.x == 0) d.sc->cands = counter[1];
    }
}

}  // namespace

size_t exact_scan_temp_bytes(u64 span)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const u32*)nullptr, (u32*)nullptr, (size_t)span);
    return bytes;
}

// list / slot / bucket: cap entries each; cnt / cursor: span entries each; counter: two words (k_exact_phase1)
cudaError_t launch_exact_phase1(const KParams& kp, const DevPtrs& d, u64* list, u32* slot, u32* cnt, u32 lo, u32 span, u32* counter,
                                int sm_count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(counter, 0, 2 * sizeof(u32), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(cnt, 0, (size_t)span * sizeof(u32), st);
    if (e != cudaSuccess || !kp.count) return e;
    u64 blocks = (kp.count + 255) / 256;
    const u64 cap = (u64)sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_exact_phase1<<<(unsigned)blocks, 256, 0, st>>>(kp, d, list, slot, cnt, lo, counter);
    return cudaGetLastError();
}
cudaError_t launch_exact_group(const u64* list, const u32* slot, const u32* counter, const u32* cnt, u32* cursor, u32 lo, u32 span,
                               u32* bucket, void* tmp, size_t tmp_bytes, int sm_count, cudaStream_t st)
{
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, cursor, (size_t)span, st);
    if (e != cudaSuccess) return e;
    k_exact_scatter<<<(unsigned)sm_count * 8, 256, 0, st>>>(list, slot, counter, cursor, lo, bucket);
    return cudaGetLastError();
}
cudaError_t launch_exact_phase3(const KParams& kp, const DevPtrs& d, u32* bucket, const u32* cnt, const u32* cursor, u32 lo, u32 span,
                                const u32* counter, int sm_count, cudaStream_t st)
{
    u64 blocks = ((u64)span + 255) / 256;
    if (blocks == 0) blocks = 1;
    const u64 cap = (u64)sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_exact_phase3<<<(unsigned)blocks, 256, 0, st>>>(kp, d, bucket, cnt, cursor, lo, span, counter);
    return cudaGetLastError();
}

}  // namespace abnn
