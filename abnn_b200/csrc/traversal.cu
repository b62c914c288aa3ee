// abnn_b200/csrc/traversal.cu — the Monte-Carlo traversal kernels (the hot path).
//
// Replaces the reference's Metal kernel monte_carlo_traversal (abnn/src/core/kernels/brain.metal:41-130)
// with the semantics fixed in DESIGN.md §2: Philox-keyed random-synapse sampling
// (README.md:77), uint64 timestamps with a per-event clock (README.md:62-63,85), lastVisited writes
// (README.md:84), and the per-event arithmetic of brain.metal:70-126 unchanged.
//
//   k_traverse_parallel : the throughput kernel. HBM-bound random 16-byte gathers of SynapsePacked,
//                         one L2 read of lastFired[src] and one L2 RED.MAX on lastVisited[dst] per
//                         event; everything past the pre-spike window gate is a rare path where
//                         same-dst lanes of a warp are serialised with match.any and timestamps
//                         move with 64-bit atomicMax.
//   k_traverse_serial   : one thread walks the events in index order (the bit-exact order of the
//                         oracle). Reference for the EXACT mode and the parity tests.
//   k_end_pass          : r-bar step, clock advance, counters -> stats slot (no host round trip).
#include "common.cuh"
#include "kernels.h"

namespace abnn {

// ---- per-event arithmetic shared by every execution mode (brain.metal:91-121) ------------------
struct Decision { float w; bool fired; };

__device__ __forceinline__ bool release_test(const KParams& kp, float w, float u)
{
    const float p = clampf(w * w * kp.base_scale, 0.f, 1.f);     // brain.metal:91
    return p > u;                                                // brain.metal:92
}
__device__ __forceinline__ float plasticity(const KParams& kp, float w, bool fired, float R, float rbar, u64 now, u64 ld)
{
    float dW = fired ? (kp.a_ltp * (1.f - w)) : (-kp.a_ltd * w);              // brain.metal:101-102
    dW += kp.eta_reward * (R - rbar) * (fired ? 1.0f : 0.0f);                 // brain.metal:107
    const float isi = (float)(now - ld);                                      // brain.metal:116
    const float est = isi > 0.f ? kp.home_tick_hz / isi : 0.f;                // brain.metal:117
    dW += kp.eta_home * (kp.target_rate_hz - est) * w;                        // brain.metal:118
    return clampf(w + dW, kp.w_min, kp.w_max);                                // brain.metal:121
}
__device__ __forceinline__ u64 event_now(const KParams& kp, u64 clock, u64 i)
{
    return kp.clock_mode == ABNN_CLOCK_PER_PASS ? clock : clock + i * kp.world + kp.rank;
}
__device__ __forceinline__ Philox4 event_philox(const KParams& kp, u64 eid)
{
    return philox4x32_10((u32)eid, (u32)(eid >> 32), kp.rank, STREAM_EVENT, kp.seed_lo, kp.seed_hi);
}
__device__ __forceinline__ void stage_growth(const KParams& kp, const DevPtrs& d, u64 eid, u64 order, u32 src, u32 trial)
{
    // README.md:125 "rand() < p_new on fire -> append (src, dst') with w_init"
    if (!(kp.p_new > 0.f) || !((float)trial * (1.0f / 4294967296.0f) < kp.p_new)) return;
    const Philox4 g = philox4x32_10((u32)eid, (u32)(eid >> 32), kp.rank, STREAM_GROW, kp.seed_lo, kp.seed_hi);
    const u32 nd = (u32)(kp.n_input + mulhi64(((u64)g.x << 32) | g.y, kp.n_neuron - kp.n_input));
    const u32 slot = atomicAdd(&d.sc->grow_count, 1u);
    if (slot < kp.grow_cap) d.grow[slot] = GrowCand{order, src, nd};
    else atomicAdd(&d.sc->grow_overflow, 1u);
    atomicAdd(&d.sc->grown_pass, 1ull);
}

// ================================================================================================
// SERIAL: strict event order, one thread. Mirrors oracle/oracle_b.cpp:ob_run_pass line for line.
__global__ void k_traverse_serial(const __grid_constant__ KParams kp, const DevPtrs d)
{
    if (blockIdx.x || threadIdx.x) return;
    DevScalars* sc = d.sc;
    const u64 clock = sc->clock, event_base = sc->event_base, tick_base = sc->tick_base;
    const float R = sc->reward;
    float rbar = sc->rbar;
    u32 fires = 0;
    u64 gated = 0, fired_n = 0, cands = 0;
    const bool need_philox = kp.sampler == ABNN_SAMPLER_PHILOX || kp.release_rng == ABNN_RNG_PHILOX || kp.p_new > 0.f;
    for (u64 i = 0; i < kp.count; ++i) {
        const u64 eid = event_base + i;
        Philox4 r{0, 0, 0, 0};
        if (need_philox) r = event_philox(kp, eid);
        u64 edge;
        if (kp.sampler == ABNN_SAMPLER_SWEEP) { edge = i; if (edge >= kp.n_local) continue; }
        else { if (!kp.n_local) break; edge = mulhi64(((u64)r.x << 32) | r.y, kp.n_local); }
        const u64 now = event_now(kp, clock, i);
        const abnn_synapse s = d.syn[edge];
        if (kp.track_visits && d.visited[s.dst] < now) d.visited[s.dst] = now;
        const u64 lp = d.view[s.src];
        if (now - lp > kp.window_pre) continue;
        ++cands;
        const u64 ld = d.live[s.dst];
        if (now - ld <= kp.refractory) continue;
        if (kp.budget_on && fires >= kp.budget_share) continue;
        const float u = kp.release_rng == ABNN_RNG_XORSHIFT ? rand01_xorshift((u32)i ^ (u32)now) : u01_24(r.z);
        const bool fired = release_test(kp, s.w, u);
        if (fired) ++fires;
        const float w = plasticity(kp, s.w, fired, R, rbar, now, ld);
        if (kp.rbar_mode == ABNN_RBAR_METAL_TID0 && i == 0 && kp.rank == 0)
            rbar = rbar + kp.alpha_rbar * (R - rbar);                         // brain.metal:110-113
        d.syn[edge].w = w;                                                    // brain.metal:122
        ++gated;
        if (fired) {
            if (d.live[s.dst] < now) d.live[s.dst] = now;                     // brain.metal:125-126
            ++fired_n;
            stage_growth(kp, d, eid, tick_base + i * kp.world + kp.rank, s.src, r.w);
        }
    }
    sc->rbar = rbar;
    sc->gated += gated; sc->fired += fired_n; sc->cands += cands; sc->fires_claimed = fires;
}

// ================================================================================================
// PARALLEL rare path: the event passed the pre-spike window. Lanes of the warp that hit the same
// destination are serialised in lane (= event) order; everything else proceeds concurrently.
__device__ __noinline__ void rare_path(const KParams& kp, const DevPtrs& d, u64 i, u64 edge, u32 src, u32 dst,
                                       float w_loaded, u64 now)
{
    const unsigned act   = __activemask();
    const unsigned peers = __match_any_sync(act, dst);
    const int my_turn    = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
    const int turns      = __reduce_max_sync(act, __popc(peers));
    DevScalars* sc = d.sc;
    const float R = sc->reward, rbar = sc->rbar;        // constant during a pass (PASS_STEP r-bar)
    bool gated = false, fired = false;
    for (int t = 0; t < turns; ++t) {
        if (t == my_turn) {
            const u64 ld = *(volatile u64*)(d.live + dst);
            bool pass = !(now - ld <= kp.refractory);                                       // brain.metal:79-83
            if (pass && kp.budget_on && *(volatile u32*)&sc->fires_claimed >= kp.budget_share) pass = false;   // :85-88
            if (pass) {
                const float w0 = t == 0 ? w_loaded : *(volatile float*)&d.syn[edge].w;      // same-edge peers see the update
                const u64 eid = sc->event_base + i;
                Philox4 r{0, 0, 0, 0};
                if (kp.release_rng == ABNN_RNG_PHILOX || kp.p_new > 0.f) r = event_philox(kp, eid);
                const float u = kp.release_rng == ABNN_RNG_XORSHIFT ? rand01_xorshift((u32)i ^ (u32)now) : u01_24(r.z);
                fired = release_test(kp, w0, u);
                if (fired && kp.budget_on) {                                                // brain.metal:95-98, saturating
                    const u32 old = atomicAdd(&sc->fires_claimed, 1u);
                    if (old >= kp.budget_share) { fired = false; atomicSub(&sc->fires_claimed, 1u); }
                }
                const float w1 = plasticity(kp, w0, fired, R, rbar, now, ld);
                *(volatile float*)&d.syn[edge].w = w1;                                      // brain.metal:122
                gated = true;
                if (fired) {
                    atomicMax(d.live + dst, now);                                           // brain.metal:125-126
                    stage_growth(kp, d, eid, sc->tick_base + i * kp.world + kp.rank, src, r.w);
                }
            }
        }
        __syncwarp(act);
    }
    const unsigned ng = __popc(__ballot_sync(act, gated)), nf = __popc(__ballot_sync(act, fired)), nc = __popc(act);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(act) - 1)) {
        atomicAdd(&sc->cands, (u64)nc);
        if (ng) atomicAdd(&sc->gated, (u64)ng);
        if (nf) atomicAdd(&sc->fired, (u64)nf);
    }
}

// 16-byte streaming gather of one SynapsePacked: L1 no-allocate, evict-first in L2 so the synapse
// stream does not displace the timestamp arrays (which are pinned with an access-policy window).
__device__ __forceinline__ uint4 load_synapse(const abnn_synapse* p)
{
    return __ldcs(reinterpret_cast<const uint4*>(p));
}

// PARALLEL: each CTA walks tiles of 256*U events; a thread keeps U independent gathers in flight.
template <int SAMPLER, int VISITS, int U>
__global__ void __launch_bounds__(256) k_traverse_parallel(const __grid_constant__ KParams kp, const DevPtrs d)
{
    const u64 clock = d.sc->clock, event_base = d.sc->event_base;
    const u64 tile = 256ull * U;
    for (u64 base = (u64)blockIdx.x * tile; base < kp.count; base += (u64)gridDim.x * tile) {
        u64   edge[U];
        uint4 s[U];
        u64   lp[U];
        bool  ok[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const u64 i = base + (u64)j * 256 + threadIdx.x;
            ok[j] = i < kp.count;
            if (SAMPLER == ABNN_SAMPLER_PHILOX) {
                const Philox4 r = event_philox(kp, event_base + i);
                edge[j] = mulhi64(((u64)r.x << 32) | r.y, kp.n_local);
            } else {
                edge[j] = i;
                ok[j] = ok[j] && i < kp.n_local;                                 // brain.metal:61
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (ok[j]) s[j] = load_synapse(d.syn + edge[j]);                     // brain.metal:70
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (ok[j]) lp[j] = __ldcg(d.view + s[j].x);                          // brain.metal:73
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (!ok[j]) continue;
            const u64 i = base + (u64)j * 256 + threadIdx.x;
            const u64 now = event_now(kp, clock, i);
            if (VISITS) atomicMax(d.visited + s[j].y, now);                      // README.md:84 (RED.MAX.64 at L2)
            if (now - lp[j] <= kp.window_pre)                                    // brain.metal:74
                rare_path(kp, d, i, edge[j], s[j].x, s[j].y, __uint_as_float(s[j].z), now);
        }
    }
}

// End of pass: r-bar EWMA (once per pass), clock (brain.metal:129 / README.md:85),
// counters into the stats slot, counters reset for the next pass.
__global__ void k_end_pass(const __grid_constant__ KParams kp, DevScalars* sc, abnn_pass_stats* out)
{
    if (blockIdx.x || threadIdx.x) return;
    if (kp.rbar_mode == ABNN_RBAR_PASS_STEP) sc->rbar = sc->rbar + kp.alpha_rbar * (sc->reward - sc->rbar);
    if (kp.clock_mode == ABNN_CLOCK_PER_PASS) { sc->clock += 1; sc->last_pass_ticks = 1; }
    else { sc->clock += kp.ticks; sc->last_pass_ticks = kp.ticks; }
    sc->tick_base += kp.ticks;
    sc->event_base += kp.max_count;
    sc->pass_index += 1;
    out->events = kp.count; out->gated = sc->gated; out->fired = sc->fired; out->candidates = sc->cands;
    out->grown = sc->grown_pass; out->clock = sc->clock; out->device_ms = 0.0; out->traverse_ms = 0.0;
    sc->gated = 0; sc->fired = 0; sc->cands = 0; sc->grown_pass = 0; sc->fires_claimed = 0;
}

// ---- launchers -----------------------------------------------------------------------------------
template <int SAMPLER, int VISITS>
static cudaError_t launch_parallel_t(const KParams& kp, const DevPtrs& d, int sm_count, cudaStream_t st)
{
    constexpr int U = 4;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_traverse_parallel<SAMPLER, VISITS, U>, 256, 0);
    if (per_sm < 1) per_sm = 1;
    const u64 tiles = (kp.count + 256ull * U - 1) / (256ull * U);
    u64 grid = (u64)sm_count * per_sm;                 // persistent: a whole number of waves
    if (grid > tiles) grid = tiles;
    if (grid == 0) return cudaSuccess;
    k_traverse_parallel<SAMPLER, VISITS, U><<<(unsigned)grid, 256, 0, st>>>(kp, d);
    return cudaGetLastError();
}

cudaError_t launch_traverse_parallel(const KParams& kp, const DevPtrs& d, int sm_count, cudaStream_t st)
{
    const bool ph = kp.sampler == ABNN_SAMPLER_PHILOX, vis = kp.track_visits != 0;
    if (ph && vis)  return launch_parallel_t<ABNN_SAMPLER_PHILOX, 1>(kp, d, sm_count, st);
    if (ph && !vis) return launch_parallel_t<ABNN_SAMPLER_PHILOX, 0>(kp, d, sm_count, st);
    if (!ph && vis) return launch_parallel_t<ABNN_SAMPLER_SWEEP, 1>(kp, d, sm_count, st);
    return launch_parallel_t<ABNN_SAMPLER_SWEEP, 0>(kp, d, sm_count, st);
}
cudaError_t launch_traverse_serial(const KParams& kp, const DevPtrs& d, cudaStream_t st)
{
    k_traverse_serial<<<1, 32, 0, st>>>(kp, d);
    return cudaGetLastError();
}
cudaError_t launch_end_pass(const KParams& kp, DevScalars* sc, abnn_pass_stats* out, cudaStream_t st)
{
    k_end_pass<<<1, 32, 0, st>>>(kp, sc, out);
    return cudaGetLastError();
}

}  // namespace abnn
