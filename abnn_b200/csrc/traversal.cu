// abnn_b200/csrc/traversal.cu — the Monte-Carlo traversal kernels (the hot path).
//
// Replaces the reference's Metal kernel monte_carlo_traversal (abnn/src/core/kernels/brain.metal:41-130)
// with the semantics fixed in DESIGN.md §2: Philox-keyed random-synapse sampling
// (README.md:77), uint64 timestamps with a per-event clock (README.md:62-63,85), lastVisited writes
// (README.md:84), and the per-event arithmetic of brain.metal:70-126 unchanged.
//
//   k_traverse_line     : the throughput kernel (PHILOX sampler, sample_block = 8): every warp copies chunks
//                         of 32 Philox-chosen 128-byte lines of the synapse table into a 4 KB shared-memory
//                         stage with cp.async; per event one L2 read of the 32-bit pre-spike gate word of
//                         src; lastVisited[dst] moves with one RED.MAX.64 per run of equal destinations
//                         (one per line when the table is dst-sorted); the events that pass the window
//                         and the refractory gate are compacted and resolved per destination, in event
//                         order, inside the warp (match.any + ballot chain) and publish with 64-bit
//                         atomicMax. 32 warps per SM hide the HBM latency.
//   k_build_slack       : the per-pass 32-bit gate words (window gate of the pass-start snapshot).
//   k_traverse_parallel : iid PHILOX sampler (16-byte random gathers) and the SWEEP sampler.
//   k_traverse_block    : sample_block = 2, 4, 16, 32 (register-staged).
//   k_traverse_serial   : one thread walks the events in index order (the bit-exact order of the
//                         oracle). Reference for the EXACT mode and the parity tests.
//   k_end_pass          : r-bar step, clock advance, counters -> stats slot (no host round trip).
#include <cstdlib>

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
__device__ __forceinline__ float plasticity_f(const KParams& kp, float w, bool fired, float R, float rbar, float isi)
{
    float dW = fired ? (kp.a_ltp * (1.f - w)) : (-kp.a_ltd * w);              // brain.metal:101-102
    dW += kp.eta_reward * (R - rbar) * (fired ? 1.0f : 0.0f);                 // brain.metal:107
    const float est = isi > 0.f ? kp.home_tick_hz / isi : 0.f;                // brain.metal:117
    dW += kp.eta_home * (kp.target_rate_hz - est) * w;                        // brain.metal:118
    return clampf(w + dW, kp.w_min, kp.w_max);                                // brain.metal:121
}
__device__ __forceinline__ float plasticity(const KParams& kp, float w, bool fired, float R, float rbar, u64 isi_ticks)
{
    return plasticity_f(kp, w, fired, R, rbar, (float)isi_ticks);             // brain.metal:116 (now - ld)
}
__device__ __forceinline__ u64 event_now(const KParams& kp, u64 clock, u64 i)
{
    return kp.clock_mode == ABNN_CLOCK_PER_PASS ? clock : clock + i * kp.world + kp.rank;
}
__device__ __forceinline__ Philox4 event_philox(const KParams& kp, u64 eid)
{
    return philox4x32_10((u32)eid, (u32)(eid >> 32), kp.rank, STREAM_EVENT, kp.seed_lo, kp.seed_hi);
}
// The two per-event random words (release draw, synaptogenesis trial) of local event i: taken from the Philox call of
// the event's sample group (common.cuh:release_word).
struct EventWords { u32 rel, trial; };
__device__ __forceinline__ EventWords event_words(const KParams& kp, u64 event_base, u64 i)
{
    const u32 group = kp.sampler == ABNN_SAMPLER_PHILOX ? kp.sample_block : 1u;
    const u32 lane = (u32)i & (group - 1u);
    const Philox4 q = event_philox(kp, event_base + i - lane);
    return EventWords{release_word(q.z, group, lane), trial_word(q.w, group, lane)};
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
    const u32 group = kp.sampler == ABNN_SAMPLER_PHILOX ? kp.sample_block : 1u;      // events per Philox call
    for (u64 i = 0; i < kp.count; ++i) {
        const u64 eid = event_base + i;
        const u32 lane = (u32)i & (group - 1u);
        Philox4 q{0, 0, 0, 0};
        if (need_philox) q = event_philox(kp, eid - lane);
        u64 edge;
        if (kp.sampler == ABNN_SAMPLER_SWEEP) { edge = i; if (edge >= kp.n_local) continue; }
        else if (kp.sample_block == 1) { if (!kp.n_local) break; edge = mulhi64(((u64)q.x << 32) | q.y, kp.n_local); }
        else {
            if (!kp.n_local) break;
            edge = (mulhi64(((u64)q.x << 32) | q.y, kp.n_blocks) << kp.log_block) + lane;
            if (edge >= kp.n_local) continue;
        }
        const u64 now = event_now(kp, clock, i);
        const abnn_synapse s = d.syn[edge];
        if (s.src == DEAD_SRC) continue;                                      // pruned, waiting for the next rebuild
        if (kp.track_visits && d.visited[s.dst] < now) d.visited[s.dst] = now;
        const u64 lp = d.view[s.src];
        if (now - lp > kp.window_pre) continue;
        ++cands;
        const u64 ld = d.live[s.dst];
        if (now - ld <= kp.refractory) continue;
        if (kp.budget_on && fires >= kp.budget_share) continue;
        const float u = kp.release_rng == ABNN_RNG_XORSHIFT ? rand01_xorshift((u32)i ^ (u32)now) : u01_24(release_word(q.z, group, lane));
        const bool fired = release_test(kp, s.w, u);
        if (fired) ++fires;
        const float w = plasticity(kp, s.w, fired, R, rbar, now - ld);
        if (kp.rbar_mode == ABNN_RBAR_METAL_TID0 && i == 0 && kp.rank == 0)
            rbar = rbar + kp.alpha_rbar * (R - rbar);                         // brain.metal:110-113
        d.syn[edge].w = w;                                                    // brain.metal:122
        stage_prune(kp, d, edge, w);
        ++gated;
        if (fired) {
            if (d.live[s.dst] < now) d.live[s.dst] = now;                     // brain.metal:125-126
            ++fired_n;
            stage_growth(kp, d, eid, tick_base + i * kp.world + kp.rank, s.src, trial_word(q.w, group, lane));
        }
    }
    sc->rbar = rbar;
    sc->gated += gated; sc->fired += fired_n; sc->cands += cands; sc->fires_claimed = fires;
}

// ================================================================================================
// PARALLEL execution.
//
// Hot part (every event): gather the SynapsePacked record, read lastFired[src] from the pass-start
// view (one L2 sector), RED.MAX.64 lastVisited[dst] (one L2 sector), test the pre-spike window.
// Gated path (events that pass the window, brain.metal:74): lanes of the warp that hit the same
// destination are serialised in lane (= event) order with match.any; timestamps move with 64-bit
// atomicMax, so "the latest event wins" holds whatever the execution order.
// (A variant that queued candidates in shared memory and processed them 32 at a time was measured
// slower — the kernel is bound by L2 sector operations and latency, not by issue slots; see
// profiles/r1_notes.md.)
struct PassConsts { u64 clock, event_base, tick_base; float R, rbar; };

// One candidate's turn: refractory + budget gates, release test, plasticity, timestamp write.
// Returns bit0 = gated (weight written), bit1 = fired.
__device__ __forceinline__ u32 candidate_turn(const KParams& kp, const DevPtrs& d, const PassConsts& pc, u64 i, u64 edge,
                                              u32 src, u32 dst, float w, u64 now, bool reload_w)
{
    // Unordered execution can observe a fire that is LATER than this event (ld > now); the unsigned
    // difference would wrap and look ancient. The gate is therefore symmetric: two fires of one neuron
    // are never closer than the refractory period. In event order ld <= now and this is brain.metal:79-83.
    const u64 ld = __ldcg(d.live + dst);
    const u64 gap = ld <= now ? now - ld : ld - now;
    if (gap <= kp.refractory) return 0;                                                    // brain.metal:79-83
    if (kp.budget_on && *(volatile u32*)&d.sc->fires_claimed >= kp.budget_share) return 0;  // brain.metal:85-88
    if (reload_w) w = __ldcg(&d.syn[edge].w);        // an earlier same-destination peer may have updated this record
    const u64 eid = pc.event_base + i;
    EventWords r{0, 0};
    if (kp.release_rng == ABNN_RNG_PHILOX || kp.p_new > 0.f) r = event_words(kp, pc.event_base, i);
    const float u = kp.release_rng == ABNN_RNG_XORSHIFT ? rand01_xorshift((u32)i ^ (u32)now) : u01_24(r.rel);
    bool fired = release_test(kp, w, u);
    if (fired && kp.budget_on) {                                                            // brain.metal:95-98, saturating
        const u32 old = atomicAdd(&d.sc->fires_claimed, 1u);
        if (old >= kp.budget_share) { fired = false; atomicSub(&d.sc->fires_claimed, 1u); }
    }
    const float w_new = plasticity(kp, w, fired, pc.R, pc.rbar, gap);
    __stcg(&d.syn[edge].w, w_new);                                                          // brain.metal:122
    stage_prune(kp, d, edge, w_new);
    if (!fired) return 1;
    atomicMax(d.live + dst, now);                                                           // brain.metal:125-126
    stage_growth(kp, d, eid, pc.tick_base + i * kp.world + kp.rank, src, r.trial);
    return 3;
}

__device__ __forceinline__ u32 gated_path(const KParams& kp, const DevPtrs& d, const PassConsts& pc, u64 i, u64 edge,
                                          u32 src, u32 dst, float w_loaded, u64 now)
{
    const unsigned act   = __activemask();
    const unsigned peers = __match_any_sync(act, dst);
    const unsigned lane  = threadIdx.x & 31;
    const unsigned conflicted = __ballot_sync(act, peers != (1u << lane));   // lanes that share a destination
    if (peers == (1u << lane))                                               // common case: destination unique in the warp
        return candidate_turn(kp, d, pc, i, edge, src, dst, w_loaded, now, false);
    const int my_turn = __popc(peers & ((1u << lane) - 1u));
    const int turns = __reduce_max_sync(conflicted, __popc(peers));
    u32 result = 0;
    for (int t = 0; t < turns; ++t) {                                        // one turn each, in lane order
        if (t == my_turn) result = candidate_turn(kp, d, pc, i, edge, src, dst, w_loaded, now, t > 0);
        __syncwarp(conflicted);
    }
    return result;
}

// Gated path without the spike budget (north-star profile). Called by the whole (converged) warp.
// Candidates of the warp that share a destination form a chain that must run in event (= lane)
// order, because a fire moves lastFired[dst] for the events behind it. The chain state is one
// register (ld): every candidate evaluates against the value in memory; the first candidate of a
// destination that fires is final together with everything before it, the candidates behind it
// re-evaluate against its timestamp — one round per fire in a chain, zero extra rounds when nothing
// fires (the common case), and no memory round trip between chain members. Exactly the serial
// order for the events of one warp; across warps lastFired moves with atomicMax.
// Returns bit0 = gated (weight written), bit1 = fired.
__device__ __forceinline__ u32 chain_resolve(const KParams& kp, const DevPtrs& d, const PassConsts& pc, bool cand, u64 i,
                                             u64 edge, u32 src, u32 dst, float w, u64 now, u64 ld)
{
    const unsigned cmask = __ballot_sync(0xffffffffu, cand);
    if (!cand) return 0;
    const unsigned lane  = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(cmask, dst);
    const u64 eid = pc.event_base + i;
    EventWords r{0, 0};
    if (kp.release_rng == ABNN_RNG_PHILOX || kp.p_new > 0.f) r = event_words(kp, pc.event_base, i);
    const float u = kp.release_rng == ABNN_RNG_XORSHIFT ? rand01_xorshift((u32)i ^ (u32)now) : u01_24(r.rel);
    const bool rel = release_test(kp, w, u);                                                // brain.metal:91-92
    bool settled = false, skip = true, fired = false;
    u64 gap = 0;
    do {
        if (!settled) {
            gap = ld <= now ? now - ld : ld - now;       // symmetric: a later event of another warp may have fired already
            skip = gap <= kp.refractory;                                                    // brain.metal:79-83
            fired = !skip && rel;
        }
        const unsigned F = __ballot_sync(cmask, fired && !settled) & peers;   // unsettled peers that fire under the current ld
        const int first = F ? __ffs(F) - 1 : (int)lane;
        const u64 now_first = __shfl_sync(cmask, now, first);
        if (!settled) {
            if (F == 0 || lane <= (unsigned)first) settled = true;           // nothing before me fires: my decision stands
            else if (now_first > ld) ld = now_first;                          // brain.metal:125-126 seen by the events behind it
        }
    } while (__any_sync(cmask, !settled));
    if (skip) return 0;
    const float w_new = plasticity(kp, w, fired, pc.R, pc.rbar, gap);
    __stcg(&d.syn[edge].w, w_new);                                                          // brain.metal:122
    stage_prune(kp, d, edge, w_new);
    if (!fired) return 1;
    if (kp.use_line32) atomicMax(d.fire32 + dst, (int)(now - pc.clock));                    // brain.metal:125-126 (32-bit form)
    else atomicMax(d.live + dst, now);                                                      // brain.metal:125-126
    stage_growth(kp, d, eid, pc.tick_base + i * kp.world + kp.rank, src, r.trial);
    return 3;
}
__device__ __forceinline__ u32 chain_path(const KParams& kp, const DevPtrs& d, const PassConsts& pc, bool cand, u64 i,
                                          u64 edge, u32 src, u32 dst, float w, u64 now)
{
    u64 ld = 0;
    if (cand) {
        if (kp.use_line32) {                             // 32-bit pass-relative fire word; beyond 2^30 ticks the 64-bit array is the truth
            const int fv = __ldcg(d.fire32 + dst);
            ld = (fv == FIRE32_ANCIENT || fv == FIRE32_FUTURE) ? __ldcg(d.live + dst) : pc.clock + (u64)(long long)fv;
        } else ld = __ldcg(d.live + dst);                                                   // brain.metal:79
    }
    return chain_resolve(kp, d, pc, cand, i, edge, src, dst, w, now, ld);
}

// Events that passed the pre-spike window: budgeted runs keep the turn-based path (the global spike
// budget is claimed per fire), everything else resolves its chains in registers.
__device__ __forceinline__ u32 resolve_candidates(const KParams& kp, const DevPtrs& d, const PassConsts& pc, bool cand, u64 i,
                                                  u64 edge, u32 src, u32 dst, float w, u64 now)
{
    if (!__any_sync(0xffffffffu, cand)) return 0;
    if (!kp.budget_on) return chain_path(kp, d, pc, cand, i, edge, src, dst, w, now);
    return cand ? gated_path(kp, d, pc, i, edge, src, dst, w, now) : 0u;
}

// lastVisited[dst] = max(., now) (README.md:84) once per run of equal destinations in adjacent lanes:
// `now` grows with the lane, so the last lane of a run carries the run's maximum. With a dst-sorted
// table a 128-byte line is one run -> one RED.MAX.64 per line instead of eight. Whole warp calls.
__device__ __forceinline__ void visit(const DevPtrs& d, bool ok, u32 dst, u64 now)
{
    const u32 key = ok ? dst : 0xFFFFFFFFu;
    const u32 next = __shfl_down_sync(0xffffffffu, key, 1);
    if (ok && ((threadIdx.x & 31) == 31 || next != key)) atomicMax(d.visited + dst, now);   // RED.MAX.64 at L2
}

// the same on the 32-bit visit words of a pass (1 + latest tick offset; folded into lastVisited by k_fold_prepare32)
__device__ __forceinline__ void visit32(const DevPtrs& d, bool ok, u32 dst, u32 t)
{
    const u32 key = ok ? dst : 0xFFFFFFFFu;
    const u32 next = __shfl_down_sync(0xffffffffu, key, 1);
    if (ok && ((threadIdx.x & 31) == 31 || next != key)) atomicMax(d.vis32 + dst, t + 1u);
}

__device__ __forceinline__ void flush_counters(const DevPtrs& d, u32 n_cand, u32 n_gated, u32 n_fired, u32* s_cnt)
{
    n_cand = __reduce_add_sync(0xffffffffu, n_cand);
    n_gated = __reduce_add_sync(0xffffffffu, n_gated);
    n_fired = __reduce_add_sync(0xffffffffu, n_fired);
    if ((threadIdx.x & 31) == 0) {
        if (n_cand) atomicAdd(&s_cnt[0], n_cand);
        if (n_gated) atomicAdd(&s_cnt[1], n_gated);
        if (n_fired) atomicAdd(&s_cnt[2], n_fired);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(&d.sc->cands, (u64)s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&d.sc->gated, (u64)s_cnt[1]);
        if (s_cnt[2]) atomicAdd(&d.sc->fired, (u64)s_cnt[2]);
    }
}

// 16-byte streaming gather of one SynapsePacked (evict-first: the synapse stream should not displace
// the timestamp arrays in L2).
__device__ __forceinline__ uint4 load_synapse(const abnn_synapse* p)
{
    return __ldcs(reinterpret_cast<const uint4*>(p));
}

// Pre-spike window test of the iid / block kernels (brain.metal:73-77) in two stages, so that the reads of a batch stay
// in flight together. With kp.use_slack (ABNN_IID_SLACK experiment) the 4-byte gate word of the pass (k_build_slack)
// replaces the 8-byte snapshot read, as in the line kernel: half the bytes, and the 20 MB gate array sits in the
// persisting L2 window.
__device__ __forceinline__ u64 window_word(const KParams& kp, const DevPtrs& d, u32 src)
{
    return kp.use_slack ? (u64)__ldcg(d.slack + src) : __ldcg(d.view + src);
}
__device__ __forceinline__ bool window_test(const KParams& kp, const DevPtrs& d, bool ok, u32 src, u64 word, u64 now, u64 clock)
{
    if (!ok) return false;
    if (kp.use_slack) {
        if ((u32)word != SLACK_EXACT) return (u32)(now - clock) < (u32)word;
        word = __ldcg(d.view + src);                                     // snapshot in the future of the pass start (rare)
    }
    return now - word <= kp.window_pre || (!kp.snapshot && word > now);  // live view: a later spike is recent
}

#ifndef ABNN_TRAV_MIN_CTAS
#define ABNN_TRAV_MIN_CTAS 3
#endif

// iid PHILOX sampler (sample_block = 1) and SWEEP sampler: each CTA walks tiles of 256*U events; a
// thread keeps U independent gathers in flight. Pass counters stay in registers and are reduced once
// per CTA (no per-event global atomics).
template <int SAMPLER, int VISITS, int U>
__global__ void __launch_bounds__(256, ABNN_TRAV_MIN_CTAS) k_traverse_parallel(const __grid_constant__ KParams kp, const DevPtrs d)
{
    __shared__ u32 s_cnt[3];
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const PassConsts pc{d.sc->clock, d.sc->event_base, d.sc->tick_base, d.sc->reward, d.sc->rbar};
    u32 n_cand = 0, n_gated = 0, n_fired = 0;
    const u64 tile = 256ull * U;
    for (u64 base = (u64)blockIdx.x * tile; base < kp.count; base += (u64)gridDim.x * tile) {
        u64   edge[U];
        uint4 s[U] = {};
        u64   lp[U];
        bool  ok[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const u64 i = base + (u64)j * 256 + threadIdx.x;
            ok[j] = i < kp.count;
            if (SAMPLER == ABNN_SAMPLER_PHILOX) {
                const Philox4 r = event_philox(kp, pc.event_base + i);
                edge[j] = mulhi64(((u64)r.x << 32) | r.y, kp.n_local);
            } else {
                edge[j] = i;
                ok[j] = ok[j] && i < kp.n_local;                                 // brain.metal:61
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (ok[j]) { s[j] = load_synapse(d.syn + edge[j]); ok[j] = s[j].x != DEAD_SRC; }   // brain.metal:70
#pragma unroll
        for (int j = 0; j < U; ++j)
            if (ok[j]) lp[j] = window_word(kp, d, s[j].x);                       // brain.metal:73
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const u64 i = base + (u64)j * 256 + threadIdx.x;
            const u64 now = event_now(kp, pc.clock, i);
            if (VISITS) { if (kp.use_line32) visit32(d, ok[j], s[j].y, (u32)(now - pc.clock)); else visit(d, ok[j], s[j].y, now); }   // README.md:84
            const bool cand = window_test(kp, d, ok[j], s[j].x, lp[j], now, pc.clock);           // brain.metal:74
            const u32 r = resolve_candidates(kp, d, pc, cand, i, edge[j], s[j].x, s[j].y, __uint_as_float(s[j].z), now);
            n_cand += cand; n_gated += r & 1u; n_fired += r >> 1;
        }
    }
    flush_counters(d, n_cand, n_gated, n_fired, s_cnt);
}

// Block sampler (sample_block = B = 2^LOGB > 1): one Philox draw selects a block-aligned run of B
// consecutive SynapsePacked records and B consecutive events process it. With B = 8 a draw is exactly
// one 128-byte HBM line — the unit B200 fetches on every L2 miss whatever the load asks for — so all 8
// records of a fetched line do work instead of 1 (profiles/r1_notes.md §1).
// A warp owns chunks of 32 groups: lane L draws the block of group L, then in iteration k the warp
// processes groups k*(32/B).. with B lanes per group reading the B records (coalesced: one line per B
// lanes); the block base travels by shuffle. One Philox per B events instead of one per event.
template <int LOGB, int VISITS>
__global__ void __launch_bounds__(256, ABNN_TRAV_MIN_CTAS) k_traverse_block(const __grid_constant__ KParams kp, const DevPtrs d)
{
    constexpr int B = 1 << LOGB;
    constexpr int GPI = 32 / B;            // groups served per warp iteration
    constexpr int KB = B < 4 ? B : 4;      // iterations batched for memory-level parallelism
    __shared__ u32 s_cnt[3];
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const PassConsts pc{d.sc->clock, d.sc->event_base, d.sc->tick_base, d.sc->reward, d.sc->rbar};
    u32 n_cand = 0, n_gated = 0, n_fired = 0;
    const unsigned lane = threadIdx.x & 31;
    const u32 rec = lane & (B - 1);
    const u64 warps_total = (u64)gridDim.x * 8, warp_global = (u64)blockIdx.x * 8 + (threadIdx.x >> 5);
    const u64 n_chunks = (kp.count + 32ull * B - 1) / (32ull * B);
    for (u64 c = warp_global; c < n_chunks; c += warps_total) {
        const u64 i0 = (c * 32 + lane) << LOGB;                  // first event of this lane's group
        u64 be = ~0ull;
        if (i0 < kp.count) {
            const Philox4 r = event_philox(kp, pc.event_base + i0);
            be = mulhi64(((u64)r.x << 32) | r.y, kp.n_blocks) << LOGB;
        }
#pragma unroll
        for (int k0 = 0; k0 < B; k0 += KB) {
            u64   ev[KB], ed[KB], lp[KB];
            uint4 s[KB] = {};
            bool  ok[KB];
#pragma unroll
            for (int kk = 0; kk < KB; ++kk) {
                const int src_lane = (k0 + kk) * GPI + (int)(lane >> LOGB);
                const u64 b = __shfl_sync(0xffffffffu, be, src_lane);
                ev[kk] = ((c * 32 + src_lane) << LOGB) + rec;
                ed[kk] = b + rec;
                ok[kk] = b != ~0ull && ev[kk] < kp.count && ed[kk] < kp.n_local;
                if (ok[kk]) { s[kk] = load_synapse(d.syn + ed[kk]); ok[kk] = s[kk].x != DEAD_SRC; }
            }
#pragma unroll
            for (int kk = 0; kk < KB; ++kk)
                if (ok[kk]) lp[kk] = window_word(kp, d, s[kk].x);
#pragma unroll
            for (int kk = 0; kk < KB; ++kk) {
                const u64 now = event_now(kp, pc.clock, ev[kk]);
                if (VISITS) { if (kp.use_line32) visit32(d, ok[kk], s[kk].y, (u32)(now - pc.clock)); else visit(d, ok[kk], s[kk].y, now); }
                const bool cand = window_test(kp, d, ok[kk], s[kk].x, lp[kk], now, pc.clock);
                const u32 r = resolve_candidates(kp, d, pc, cand, ev[kk], ed[kk], s[kk].x, s[kk].y, __uint_as_float(s[kk].z), now);
                n_cand += cand; n_gated += r & 1u; n_fired += r >> 1;
            }
        }
    }
    flush_counters(d, n_cand, n_gated, n_fired, s_cnt);
}


// ================================================================================================
// Pre-spike window gate in 32 bits. With the SNAPSHOT src view the gate of event e (brain.metal:73-77)
//     now(e) - lastFired_snapshot[src] <= window_pre,      now(e) = clock + t(e),  0 <= t < ticks
// depends on src only through  v[src] = window_pre - (clock - lastFired_snapshot[src]) + 1 :
// the event passes iff t(e) < v[src]. v is rebuilt before every pass (one streaming sweep over the
// snapshot, 12 bytes per neuron) and clamped to 32 bits: 0 = never passes in this pass, >= ticks =
// always passes. The 5M-neuron gate array that every event reads is then 20 MB instead of 40 MB, which
// is what lets it stay L2-resident next to lastVisited/lastFired under the 1B-synapse stream.
// 0xFFFFFFFF marks the (pathological) neurons whose snapshot lies in the future of the pass start;
// for those the kernel falls back to the exact 64-bit test.
__global__ void __launch_bounds__(256) k_build_slack(const __grid_constant__ KParams kp, const DevPtrs d, const u64* src, u64 n0, u64 n1)
{
    const u64 clock = d.sc->clock;
    for (u64 n = n0 + (u64)blockIdx.x * blockDim.x + threadIdx.x; n < n1; n += (u64)gridDim.x * blockDim.x) {
        d.slack[n] = slack_word(clock, src[n], kp.window_pre);
    }
}
cudaError_t launch_build_slack(const KParams& kp, const DevPtrs& d, const u64* src, u64 n0, u64 n1, cudaStream_t st)
{
    if (n1 <= n0) return cudaSuccess;
    u64 blocks = (n1 - n0 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_build_slack<<<(unsigned)blocks, 256, 0, st>>>(kp, d, src, n0, n1);
    return cudaGetLastError();
}

// ================================================================================================
// Line sampler (sample_block = 8): one Philox draw = one 128-byte line of the table = 8 events.
//
// A warp walks chunks of 32 lines = 256 events. Lane L draws the line of group L; the warp copies the
// chunk into its 4 KB shared-memory stage with 8 cp.async (LDGSTS.128) instructions, each moving 4 whole
// lines (8 lanes x 16 B per line, fully coalesced), draws the lines of its NEXT chunk while the copy is
// in flight, and then consumes the stage in three phases:
//   A  8 steps; in step k lane l holds record (l & 7) of group 4k + (l >> 3), i.e. the warp reads 512
//      contiguous bytes of the stage. The pre-spike gate word of src (32-bit slack, k_build_slack) and
//      lastFired[dst] of ABNN_LINE_PART = 4 steps are requested back to back (8 independent L2 reads per
//      lane; all 8 steps at once needs 80 registers and costs the fourth CTA per SM), then B runs on those
//      steps, then the other four.
//   B  8 steps: pre-spike window test (brain.metal:74), lastVisited RED once per run of equal
//      destinations, refractory gate against the value read in A (brain.metal:79): the events that are
//      still open are COMPACTED — their 8-bit chunk-local index goes to a shared-memory queue in event
//      order. (In an active network most events pass the window and most of those are refractory; a fire
//      only moves lastFired[dst] forward, so an event that is refractory against memory stays refractory.)
//   C  ceil(open/32) dense steps over the queue: refractory gate again (now also against the chunk's own
//      fires), release draw, plasticity, weight write-back, fire. The expensive path runs with full warps
//      instead of with the ~quarter of lanes that are open in a raw step; lastFired[dst] of step j+1 is
//      fetched before step j is resolved. Same-destination events are ordered inside a step by
//      chain_resolve and across the steps of a chunk by a small shared-memory list of the chunk's
//      fires, so a warp's 256 events resolve exactly as in the serial order — with one exception in THIS kernel: two
//      groups of a chunk that drew the same line and whose events land in the same dense step both use the weight read
//      before either wrote (tables of a few hundred lines only; k_traverse_line32 cuts the step instead).
//
// Shared memory is kept to 4.5 KB per warp on purpose. Measured on B200 (profiles/r1_notes.md): the
// gathers of this kernel are limited by the L1's capacity to track outstanding misses, i.e. by what the
// shared-memory carve-out leaves of the 228 KB array. A 3-deep ring (13.5 KB/warp, 16 warps) ran at
// 3.4 ms/pass, a 2-deep ring 2.1 ms, this single stage with 24 warps 1.7 ms and with 32 warps (63
// registers) 1.46 ms; forcing the carve-out to 100 % shared memory took the same code from 2.1 to 3.9 ms.
// HBM latency is hidden across warps.
// (Also measured and dropped: one cp.async.bulk (TMA) + mbarrier per line with an L2 evict_first
// policy — 7 % slower than LDGSTS at equal depth.)
#ifndef ABNN_LINE_MIN_CTAS
#define ABNN_LINE_MIN_CTAS 4
#endif
#ifndef ABNN_LINE_PART
#define ABNN_LINE_PART 4                 // steps whose gate / lastFired reads are in flight together (8, 4 or 2)
#endif
constexpr int LINE_STAGE_BYTES = 32 * 128;
#ifndef ABNN_LINE_WARPS
#define ABNN_LINE_WARPS 8
#endif
constexpr int LINE_WARPS = ABNN_LINE_WARPS;
constexpr u32 LINE_FIRE_CAP = 16;       // fires of one chunk kept in shared memory for the chunk's later steps
constexpr size_t LINE_WARP_SMEM = LINE_STAGE_BYTES + 256 + LINE_FIRE_CAP * (sizeof(u64) + sizeof(u32)) + 64;   // 4608
constexpr size_t LINE_SMEM = LINE_WARPS * LINE_WARP_SMEM;

__device__ __forceinline__ void cp_async16(u32 dst_smem, const void* src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int VISITS, int SLACK>
__global__ void __launch_bounds__(LINE_WARPS * 32, ABNN_LINE_MIN_CTAS) k_traverse_line(const __grid_constant__ KParams kp, const DevPtrs d)
{
    constexpr int LOGB = 3, B = 8, PART = ABNN_LINE_PART;
    extern __shared__ __align__(128) unsigned char line_smem[];
    __shared__ u32 s_cnt[3];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-warp shared memory: stage | candidate queue | fire list
    unsigned char* stage = line_smem + warp * LINE_WARP_SMEM;
    unsigned char* queue = stage + LINE_STAGE_BYTES;
    u64* fl_now = reinterpret_cast<u64*>(queue + 256);
    u32* fl_dst = reinterpret_cast<u32*>(fl_now + LINE_FIRE_CAP);
    const unsigned char* mine = stage + lane * 16;
    const u32 mine_addr = (u32)__cvta_generic_to_shared(mine);
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const PassConsts pc{d.sc->clock, d.sc->event_base, d.sc->tick_base, d.sc->reward, d.sc->rbar};
    u32 n_cand = 0, n_gated = 0, n_fired = 0;
    const u32 rec = lane & (B - 1), sub = lane >> LOGB;
    const unsigned lt = (1u << lane) - 1u;
    const u64 n_chunks = (kp.count + 32ull * B - 1) / (32ull * B);
    const u32 tick = kp.clock_mode != ABNN_CLOCK_PER_PASS ? kp.world : 0u;   // now(event i) = clock + i * world + rank | clock

    // lane L draws the line of group L of chunk c: line base (a multiple of 8) | (valid records - 1), ~0 = none
    auto draw = [&](u64 c) -> u64 {
        const u64 i0 = (c * 32 + lane) << LOGB;
        if (c >= n_chunks || i0 >= kp.count) return ~0ull;
        const Philox4 r = event_philox(kp, pc.event_base + i0);
        const u64 be = mulhi64(((u64)r.x << 32) | r.y, kp.n_blocks) << LOGB;
        u64 valid = kp.n_local - be;                         // the table's last line may be short,
        if (kp.count - i0 < valid) valid = kp.count - i0;    // and so may the pass's last group
        return be | ((valid < B ? valid : B) - 1);
    };

    // Work distribution: the first 13/16 of the chunks go round-robin (no shared state), the rest are handed
    // out one by one by a device-wide ticket so that warps that drew cheap chunks take more of the tail.
    // (Tickets for every chunk cost 25 % in the read-dominated regime: one contended L2 address.)
    const u64 warps_total = (u64)gridDim.x * LINE_WARPS, warp_global = (u64)blockIdx.x * LINE_WARPS + warp;
    const u64 static_rounds = (n_chunks / warps_total) * 13 / 16;
    u64 round = 0;
    auto take = [&]() -> u64 {
        if (round < static_rounds) return warp_global + (round++) * warps_total;
        u32 t = 0;
        if (lane == 0) t = atomicAdd(&d.sc->chunk_ticket, 1u);
        return static_rounds * warps_total + __shfl_sync(0xffffffffu, t, 0);
    };
    u64 c = take();
    u64 m = draw(c);
    while (c < n_chunks) {
        // ---- stage the chunk ------------------------------------------------------------------------
        u32 okm = 0;
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const u64 mk = __shfl_sync(0xffffffffu, m, k * 4 + sub);
            if ((u32)(mk >> 32) != 0xFFFFFFFFu && rec <= ((u32)mk & 7u)) {
                okm |= 1u << k;
                cp_async16(mine_addr + k * 512, d.syn + (mk & ~7ull) + rec);
            }
        }
        cp_async_commit();
        // two groups of the chunk drew the same line (small tables only): the later one must see the
        // weights the earlier one wrote -> it re-reads them after a fence (bit g = group g repeats a line)
        const unsigned same = __match_any_sync(0xffffffffu, m);
        const unsigned dupm = __ballot_sync(0xffffffffu, m != ~0ull && (same & lt) != 0);
        const u64 c_next = take();
        const u64 m_next = draw(c_next);                     // ALU work under the copy's latency
        const u64 ev0 = c * (32ull * B);                     // first event of the chunk
        const u64 now0 = tick ? pc.clock + ev0 * kp.world + kp.rank : pc.clock;
        const u32 t0 = (u32)(now0 - pc.clock);               // SLACK: ticks fit 32 bits (abnn_run_pass)
        cp_async_wait<0>();
        __syncwarp();                                        // the other lanes' copies have landed too

        // ---- A + B, ABNN_LINE_PARTS steps at a time (fewer live registers -> 4 CTAs per SM) ----------------
        u32 candm = 0, nC = 0;
#pragma unroll
        for (int k0 = 0; k0 < B; k0 += PART) {
            // A: gate and lastFired[dst] reads in flight
            u64 ts[PART];                                    // lastFired[dst] (SLACK) / lastFired[src], then [dst]
            u32 gate[PART];                                  // SLACK: v[src]
#pragma unroll
            for (int j = 0; j < PART; ++j) {
                const int k = k0 + j;
                ts[j] = 0; gate[j] = 0;
                if ((okm >> k) & 1u) {
                    const uint2 sd = *reinterpret_cast<const uint2*>(mine + k * 512);
                    if (sd.x == DEAD_SRC) okm &= ~(1u << k);                                // pruned, waiting for the next rebuild
                    else if (SLACK) {
                        gate[j] = __ldcg(d.slack + sd.x);                                   // brain.metal:73 (32-bit form)
                        ts[j] = __ldcg(d.live + sd.y);                                      // brain.metal:79
                    } else {
                        ts[j] = __ldcg(d.view + sd.x);                                      // brain.metal:73
                    }
                }
            }
            if (SLACK) {
                u32 exact = 0;
#pragma unroll
                for (int j = 0; j < PART; ++j) {
                    const int k = k0 + j;
                    candm |= (u32)(((okm >> k) & 1u) && t0 + (k * 32 + lane) * tick < gate[j]) << k;   // brain.metal:74
                    exact |= (u32)(gate[j] == SLACK_EXACT) << k;
                }
                if (__any_sync(0xffffffffu, exact & okm)) {  // snapshot in the future of the pass start: exact 64-bit test
#pragma unroll
                    for (int j = 0; j < PART; ++j) {
                        const int k = k0 + j;
                        if (((exact & okm) >> k) & 1u) {
                            const u64 now = now0 + (u64)((k * 32 + lane) * tick);
                            const bool cand = now - __ldcg(d.view + *reinterpret_cast<const u32*>(mine + k * 512)) <= kp.window_pre;
                            candm = (candm & ~(1u << k)) | ((u32)cand << k);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < PART; ++j) {             // 64-bit gate: window test, then lastFired[dst] reads in flight
                    const int k = k0 + j;
                    const u64 now = now0 + (u64)((k * 32 + lane) * tick);
                    const bool cand = ((okm >> k) & 1u) && (now - ts[j] <= kp.window_pre || (!kp.snapshot && ts[j] > now));   // brain.metal:74
                    ts[j] = now;
                    if (cand) { candm |= 1u << k; ts[j] = __ldcg(d.live + *reinterpret_cast<const u32*>(mine + k * 512 + 4)); }   // brain.metal:79
                }
            }
            // B: lastVisited, refractory gate, compaction of the open events
#pragma unroll
            for (int j = 0; j < PART; ++j) {
                const int k = k0 + j;
                const u64 now = now0 + (u64)((k * 32 + lane) * tick);
                if (VISITS) visit(d, (okm >> k) & 1u, *reinterpret_cast<const u32*>(mine + k * 512 + 4), now);   // README.md:84
                const u64 gap = ts[j] <= now ? now - ts[j] : ts[j] - now;
                const bool open = ((candm >> k) & 1u) && gap > kp.refractory;               // brain.metal:79-83
                const unsigned cm = __ballot_sync(0xffffffffu, open);
                if (open) queue[nC + __popc(cm & lt)] = (unsigned char)(k * 32 + lane);
                nC += __popc(cm);
            }
        }
        n_cand += __popc(candm);
        __syncwarp();

        // ---- C: dense steps over the queue ----------------------------------------------------------------
        u32 nf = 0;                                          // fires of this chunk so far (warp-uniform)
        bool spilled = false;                                // fire list overflowed: later steps re-read lastFired
        u32 le = 0; uint4 sy = make_uint4(0, 0, 0, 0); u64 ld = 0;
        bool cand = lane < nC;
        if (cand) {
            le = queue[lane];
            sy = *reinterpret_cast<const uint4*>(stage + le * 16);                          // brain.metal:70
            ld = __ldcg(d.live + sy.y);                                                     // brain.metal:79
        }
#pragma unroll 1
        for (u32 j = 0; j < nC; j += 32) {
            u32 le_n = 0; uint4 sy_n = make_uint4(0, 0, 0, 0); u64 ld_n = 0;
            const bool cand_n = j + 32 + lane < nC;
            if (cand_n) {                                    // next step's record and lastFired[dst] on their way
                le_n = queue[j + 32 + lane];
                sy_n = *reinterpret_cast<const uint4*>(stage + le_n * 16);
                ld_n = __ldcg(d.live + sy_n.y);
            }
            const u64 edge = (__shfl_sync(0xffffffffu, m, le >> LOGB) & ~7ull) + (le & 7u);
            const u64 ev = ev0 + le;
            const u64 now = now0 + (u64)(le * tick);
            float w = __uint_as_float(sy.z);
            if (cand) {
                if ((dupm >> (le >> LOGB)) & 1u) w = __ldcg(&d.syn[edge].w);
                if (spilled) ld = __ldcg(d.live + sy.y);
                else for (u32 f = 0; f < nf; ++f) {          // fires of earlier steps of this chunk, in event order
                    const u64 fn = fl_now[f];
                    if (fl_dst[f] == sy.y && fn > ld) ld = fn;
                }
            }
            u32 r;
            if (kp.budget_on) r = cand ? gated_path(kp, d, pc, ev, edge, sy.x, sy.y, w, now) : 0u;
            else r = chain_resolve(kp, d, pc, cand, ev, edge, sy.x, sy.y, w, now, ld);
            n_gated += r & 1u; n_fired += r >> 1;
            const unsigned fm = __ballot_sync(0xffffffffu, (r >> 1) != 0);
            if (fm) {
                const u32 at = nf + __popc(fm & lt);
                if ((r >> 1) && at < LINE_FIRE_CAP) { fl_dst[at] = sy.y; fl_now[at] = now; }
                nf += __popc(fm);
                if (nf > LINE_FIRE_CAP) { nf = LINE_FIRE_CAP; spilled = true; }
                __syncwarp();
            }
            if (dupm | (u32)spilled) __threadfence();        // rare: make this step's writes visible to the re-reads
            cand = cand_n; le = le_n; sy = sy_n; ld = ld_n;
        }
        __syncwarp();                                        // every lane is done with the stage
        c = c_next; m = m_next;
    }
    flush_counters(d, n_cand, n_gated, n_fired, s_cnt);
}

// ================================================================================================
// k_traverse_line32 — the line sampler on 32-bit pass-relative timestamps (what bench.py times).
//
// Same walk as k_traverse_line (a warp copies chunks of 32 Philox-chosen 128-byte lines into its 4 KB stage, gate
// reads in flight, refractory prefilter, compaction, dense steps), but every time in the kernel is a 32-bit tick
// offset from the pass-start clock:
//   slack32[src]  pre-spike gate word (k_build_slack)                      20 MB at 5M neurons
//   fire32[dst]   lastFired - clock, signed, moved with atomicMax.s32       20 MB
//   vis32[dst]    1 + latest visit offset of the pass, RED.MAX.U32         20 MB
// 60 MB = exactly the persisting-L2 carve-out, so the three arrays every event touches stay resident under the 16 GB
// stream and the only DRAM traffic left is the table itself (the 64-bit lastFired array cost one 128-byte DRAM line per
// miss — as much as the table line of the event that caused it; profiles/r2_notes.md). k_prepare32 / k_fold32 convert
// from / to the 64-bit arrays around the pass.
// Preconditions (launch_traverse_parallel checks them, else k_traverse_line runs): per-event clock, no spike budget,
// snapshot src view with gate words, ticks of the pass < 2^30 - 1, refractory < 2^30, and refractory >= the tick span of
// one chunk (256 * world). The last one makes the in-warp ordering a one-round affair: inside a chunk, the first event
// of a destination that fires blocks every later event of that destination (they are within the refractory period of
// its tick), events before it keep the decision they took against memory — exactly what chain_resolve iterates to.
// One Philox call per line: .x.y choose the line, .z/.w carry the release / growth words of its 8 events
// (common.cuh:release_word).
// SB = sample_block: 8 (one 128-byte line per draw) or 16 (two consecutive lines = 256 bytes per draw, the size at which
// random HBM3e reads reach the copy bandwidth: 6.7 TB/s against 4.7 TB/s for single lines, profiles/r2_notes.md).
// SB = 1 (iid sampler, README.md:77): the same kernel with one draw per EVENT. A lane makes the 8 Philox calls of its 8
// events of the chunk and requests their 16-byte records with cp.async — 256 random gathers in flight per warp, 8,192 per
// SM, none of them holding a register (the register-staged iid kernel k_traverse_parallel keeps 3,072 in flight and runs
// at 64 % of the bare-gather rate, profiles/r2_notes.md §7). Phases A / B / C are unchanged; a dense step repeats the
// Philox call of its event for the record index and the release / growth words (28 % of the events get that far). On
// tables below 2^24 records every dense step checks for records drawn twice in the chunk, as above for lines.
// Shared memory of a warp: the 4 KB stage, then the records of the chunk's OPEN events compacted in event order (16 B
// each: src, dst, w, fire word). With the open events copied out, the stage is free as soon as phase B is over, and the
// NEXT chunk's lines are requested before the dense steps run: the HBM latency of chunk c+1 hides under phase C of chunk
// c (the kernel is latency-bound: 32 warps per SM, every warp alternates between waiting for its lines and working on
// them). A chunk with more than LINE32_CQ open events is handled in place (stage read by the dense steps, copy issued
// afterwards) — 96 covers the benchmark's regime (72 +- 7 open events per chunk) and keeps shared memory at 47 KB per CTA.
#ifndef ABNN_LINE_CQ
#define ABNN_LINE_CQ 0
#endif
#ifndef ABNN_LINE_EARLY
#define ABNN_LINE_EARLY 1
#endif
#ifndef ABNN_LINE_STEPS
#define ABNN_LINE_STEPS 8
#endif
// ABNN_LINE_TMA=1 (measurement builds): the stage is filled by cp.async.bulk — lane L issues ONE bulk copy of its whole
// 128-byte line, completion on a per-warp mbarrier — instead of 8 LDGSTS.128 per lane. Measured slower (profiles/r2_notes.md
// §9), kept as the evidence for the choice.
#ifndef ABNN_LINE_TMA
#define ABNN_LINE_TMA 0
#endif
constexpr u32 LINE32_CQ = ABNN_LINE_CQ;                    // 0: no compacted copy — the dense steps read the stage in place
constexpr bool LINE32_EARLY = LINE32_CQ > 0 && ABNN_LINE_EARLY;
constexpr int LINE32_STEPS = ABNN_LINE_STEPS;              // steps of 4 lines per chunk: a chunk is 4 * STEPS lines
constexpr u32 LINE32_LPC = 4 * LINE32_STEPS, LINE32_EPC = 32 * LINE32_STEPS;
constexpr size_t LINE32_STAGE = (size_t)LINE32_STEPS * 512;
constexpr size_t LINE32_WARP_SMEM = LINE32_STAGE + LINE32_CQ * 16 + 256 + LINE_FIRE_CAP * sizeof(u32);
constexpr size_t LINE32_SMEM = LINE_WARPS * LINE32_WARP_SMEM;

template <int VISITS, int GROW, int SB>
__global__ void __launch_bounds__(LINE_WARPS * 32, ABNN_LINE_MIN_CTAS) k_traverse_line32(const __grid_constant__ KParams kp, const DevPtrs d)
{
    constexpr int LOGB = 3, B = 8, PART = ABNN_LINE_PART;     // a LINE is 8 records; a sample group is LPG lines
    constexpr int LPG = SB == 16 ? 2 : 1, LOGSB = SB == 8 ? 3 : 4;
    constexpr bool IID = SB == 1;                             // one draw per EVENT: see "iid" in the header
    static_assert(SB == 1 || SB == 8 || SB == 16, "line kernel: sample_block 1, 8 or 16");
    extern __shared__ __align__(128) unsigned char line_smem[];
    __shared__ u32 s_cnt[3];
#if ABNN_LINE_TMA
    __shared__ __align__(8) u64 s_mbar[LINE_WARPS];
#endif
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NS = LINE32_STEPS;
    unsigned char* stage = line_smem + warp * LINE32_WARP_SMEM;
#if ABNN_LINE_TMA
    const u32 mbar = (u32)__cvta_generic_to_shared(&s_mbar[warp]);
    u32 tma_phase = 0;
    bool tma_pending = false;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
#endif
    uint4* cq = reinterpret_cast<uint4*>(stage + LINE32_STAGE);                // open events of the chunk: src, dst, w, fire word
    unsigned char* queue = stage + LINE32_STAGE + LINE32_CQ * 16;              // their chunk-local event indices
    u32* fl_dst = reinterpret_cast<u32*>(queue + 256);             // destinations that fired in this chunk
    const unsigned char* mine = stage + lane * 16;
    const u32 mine_addr = (u32)__cvta_generic_to_shared(mine);
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u64 clock = d.sc->clock, event_base = d.sc->event_base, tick_base = d.sc->tick_base;
    const float R = d.sc->reward, rbar = d.sc->rbar;
    u32 n_cand = 0, n_gated = 0, n_fired = 0;
    const u32 rec = lane & (B - 1), sub = lane >> LOGB;
    const unsigned lt = (1u << lane) - 1u;
    const u32 count = (u32)kp.count;                               // < 2^30 (ticks < 2^30)
    const u32 n_chunks = (count + LINE32_EPC - 1) / LINE32_EPC;
    const u32 world = kp.world, refr = (u32)kp.refractory;

    // lane L holds LINE L of chunk c: m = line base (a multiple of 8) | (valid records - 1), ~0 = none; zw = .z | .w << 32 of
    // the Philox call of the line's sample group (release / growth words of its events). With two lines per group both
    // lanes of a group evaluate the group's draw (same warp instructions, no shuffle).
    u64 zw = 0, zw_next = 0;
    auto draw = [&](u32 c, u64& zw_out) -> u64 {
        if (IID) return 0ull;                                // nothing is kept per line: stage_issue / the dense step draw per event
        const u32 i0 = (c * (LINE32_LPC / LPG) + (lane / LPG)) << LOGSB;    // first event of the line's group
        if (c >= n_chunks || lane >= LINE32_LPC || i0 >= count) return ~0ull;
        const Philox4 r = event_philox(kp, event_base + i0);
        zw_out = (u64)r.z | ((u64)r.w << 32);
        const u64 be = mulhi64(((u64)r.x << 32) | r.y, kp.n_blocks) << LOGSB;
        u64 valid = kp.n_local - be;                         // the table's last block may be short,
        if (count - i0 < valid) valid = count - i0;          // and so may the pass's last group
        const u32 off = (lane % LPG) * B;                    // this line's records inside the group
        if (valid <= off) return ~0ull;
        valid -= off;
        return (be + off) | ((valid < B ? valid : B) - 1);
    };
    const u32 warps_total = gridDim.x * LINE_WARPS, warp_global = blockIdx.x * LINE_WARPS + warp;
    const u32 static_rounds = (n_chunks / warps_total) * 13 / 16;
    u32 round = 0;
    auto take = [&]() -> u32 {
        if (round < static_rounds) return warp_global + (round++) * warps_total;
        u32 t = 0;
        if (lane == 0) t = atomicAdd(&d.sc->chunk_ticket, 1u);
        return static_rounds * warps_total + __shfl_sync(0xffffffffu, t, 0);
    };
    // request the lines of a chunk (8 LDGSTS.128 per lane, each warp instruction moves 4 whole lines); bit k of the result:
    // this lane's record of step k exists
    auto stage_issue = [&](u64 mm, u32 cc) -> u32 {
        u32 ok = 0;
        if (IID) {                                           // 8 independent Philox calls and 16-byte copies per lane
            if (cc >= n_chunks) return 0u;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const u32 i = cc * LINE32_EPC + k * 32 + lane;
                if (i < count) {
                    const Philox4 r = event_philox(kp, event_base + i);
                    ok |= 1u << k;
                    cp_async16(mine_addr + k * 512, d.syn + mulhi64(((u64)r.x << 32) | r.y, kp.n_local));
                }
            }
            cp_async_commit();
            return ok;
        }
#if ABNN_LINE_TMA
        if (!IID) {
            // lane L copies line L (stage + L * 128) in one bulk operation; the warp's mbarrier counts the bytes
            const bool have = lane < LINE32_LPC && (u32)(mm >> 32) != 0xFFFFFFFFu;
            const u32 bytes = have ? (((u32)mm & 7u) + 1u) * 16u : 0u;
            const u32 total = __reduce_add_sync(0xffffffffu, bytes);
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const u64 mk = __shfl_sync(0xffffffffu, mm, k * 4 + sub);
                if ((u32)(mk >> 32) != 0xFFFFFFFFu && rec <= ((u32)mk & 7u)) ok |= 1u << k;
            }
            tma_pending = total != 0;
            if (total) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the stage was read / written through the generic proxy
                if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(total) : "memory");
                __syncwarp();
                if (have)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"((u32)__cvta_generic_to_shared(stage + lane * 128)), "l"(d.syn + (mm & ~7ull)), "r"(bytes), "r"(mbar)
                                 : "memory");
            }
            return ok;
        }
#endif
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const u64 mk = __shfl_sync(0xffffffffu, mm, k * 4 + sub);
            if ((u32)(mk >> 32) != 0xFFFFFFFFu && rec <= ((u32)mk & 7u)) {
                ok |= 1u << k;
                cp_async16(mine_addr + k * 512, d.syn + (mk & ~7ull) + rec);
            }
        }
        cp_async_commit();
        return ok;
    };
    // bit g: line g of the chunk repeats an earlier line of the chunk (the later copy must see the weights the earlier one
    // wrote). Only tables below 2^24 lines are checked: beyond that a repeat inside a chunk has probability < 3e-5 and
    // means one event reading a weight that a concurrent event is updating — the race PARALLEL execution has between
    // warps anyway.
    const bool check_dups = kp.n_blocks < (1ull << 24);
    auto dup_mask = [&](u64 mm) -> unsigned {
        if (!check_dups) return 0u;
        if (IID) return 0xFFFFFFFFu;                         // small table: every dense step looks for repeated records
        const unsigned same = __match_any_sync(0xffffffffu, mm);
        return __ballot_sync(0xffffffffu, mm != ~0ull && (same & lt) != 0);
    };
    u32 c = take();
    u64 m = draw(c, zw);
    u32 okm = stage_issue(m, c);
    unsigned dupm = dup_mask(m);
    while (c < n_chunks) {
        const u32 c_next = take();
        const u64 m_next = draw(c_next, zw_next);            // ALU work under the copy's latency
        const u32 ev0 = c * LINE32_EPC;                      // first event of the chunk (local index)
        const u32 t0 = ev0 * world + kp.rank;                // its tick offset: now = clock + t
#if ABNN_LINE_TMA
        if (!IID) {
            if (tma_pending) {
                u32 done = 0;
                while (!done)
                    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                                 : "=r"(done) : "r"(mbar), "r"(tma_phase), "r"(0x989680u) : "memory");   // suspend-time hint: sleep, do not spin
                tma_phase ^= 1u;
            }
        } else
#endif
        cp_async_wait<0>();
        __syncwarp();

        // ---- A + B: gate words and fire32[dst] in flight, window test, visit, refractory prefilter, compaction ----
        u32 candm = 0, nC = 0;
#pragma unroll
        for (int k0 = 0; k0 < NS; k0 += PART) {
            u32 gate[PART];
            int fire[PART];
            u32 exact = 0;
#pragma unroll
            for (int j = 0; j < PART; ++j) {
                const int k = k0 + j;
                gate[j] = 0; fire[j] = 0;
                if ((okm >> k) & 1u) {
                    const uint2 sd = *reinterpret_cast<const uint2*>(mine + k * 512);
                    if (sd.x == DEAD_SRC) okm &= ~(1u << k);                                // pruned, waiting for the next rebuild
                    else {
                        gate[j] = __ldcg(d.slack + sd.x);                                   // brain.metal:73 (32-bit form)
                        fire[j] = __ldcg(d.fire32 + sd.y);                                  // brain.metal:79 (32-bit form)
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < PART; ++j) {
                const int k = k0 + j;
                candm |= (u32)(((okm >> k) & 1u) && t0 + (k * 32 + lane) * world < gate[j]) << k;   // brain.metal:74
                exact |= (u32)(gate[j] == SLACK_EXACT) << k;
            }
            if (__any_sync(0xffffffffu, exact & okm)) {      // snapshot in the future of the pass start: exact 64-bit test
#pragma unroll
                for (int j = 0; j < PART; ++j) {
                    const int k = k0 + j;
                    if (((exact & okm) >> k) & 1u) {
                        const u64 now = clock + t0 + (k * 32 + lane) * world;
                        const bool cand = now - __ldcg(d.view + *reinterpret_cast<const u32*>(mine + k * 512)) <= kp.window_pre;
                        candm = (candm & ~(1u << k)) | ((u32)cand << k);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < PART; ++j) {
                const int k = k0 + j;
                const u32 t = t0 + (k * 32 + lane) * world;
                const bool ok = (okm >> k) & 1u;
                if (VISITS) {                                                               // README.md:84, once per run of equal dst
                    const u32 key = ok ? *reinterpret_cast<const u32*>(mine + k * 512 + 4) : 0xFFFFFFFFu;
                    const u32 next = __shfl_down_sync(0xffffffffu, key, 1);
                    if (ok && (lane == 31 || next != key)) atomicMax(d.vis32 + key, t + 1u);
                }
                const int gap = (int)t - fire[j];                                           // |gap| < 2^31
                const bool open = ((candm >> k) & 1u) && (u32)(gap < 0 ? -gap : gap) > refr;   // brain.metal:79-83
                const unsigned cm = __ballot_sync(0xffffffffu, open);
                if (open) {
                    const u32 pos = nC + __popc(cm & lt);
                    queue[pos] = (unsigned char)(k * 32 + lane);
                    if (LINE32_CQ == 0) {                    // the fire word travels to the dense step in the record's unused pad slot
                        *reinterpret_cast<int*>(const_cast<unsigned char*>(mine) + k * 512 + 12) = fire[j];
                    } else if (pos < LINE32_CQ) {            // the record leaves the stage, with the fire word in its pad slot
                        uint4 r = *reinterpret_cast<const uint4*>(mine + k * 512);
                        r.w = (u32)fire[j];
                        cq[pos] = r;
                    }
                }
                nC += __popc(cm);
            }
        }
        n_cand += __popc(candm);
        __syncwarp();

        // the open events are out of the stage: the next chunk's lines can come in while the dense steps run
        const bool in_cq = LINE32_CQ > 0 && nC <= LINE32_CQ;     // the dense steps read the compacted copy
        const bool early = LINE32_EARLY && in_cq;
        u32 okm_next = 0; unsigned dupm_next = 0;
        if (early) { okm_next = stage_issue(m_next, c_next); dupm_next = dup_mask(m_next); }

        // ---- C: dense steps over the queue ----------------------------------------------------------------
        u32 nf = 0;                                          // destinations that fired in this chunk so far (warp-uniform)
        bool spilled = LINE32_CQ > 0 && !in_cq;              // overflow of the compacted copy: in place, fire word re-read
        u32 le = 0; uint4 sy = make_uint4(0, 0, 0, 0);            // sy.w = fire32[dst] as read in phase A
        bool cand = lane < nC;
        if (cand) {
            le = queue[lane];
            sy = in_cq ? cq[lane] : *reinterpret_cast<const uint4*>(stage + le * 16);      // brain.metal:70
        }
        u32 j = 0;
#pragma unroll 1
        while (j < nC) {
            u32 le_n = 0; uint4 sy_n = make_uint4(0, 0, 0, 0);
            const bool cand_n = j + 32 + lane < nC;
            if (cand_n) {                                    // next step's record
                le_n = queue[j + 32 + lane];
                sy_n = in_cq ? cq[j + 32 + lane] : *reinterpret_cast<const uint4*>(stage + le_n * 16);
            }
            u32 g = le >> LOGB, r8 = le & 7u;
            u64 edge, zwg;
            if (IID) {                                       // the event's own draw again (its record came in by cp.async)
                const Philox4 q = event_philox(kp, event_base + ev0 + le);
                edge = mulhi64(((u64)q.x << 32) | q.y, kp.n_local);
                zwg = (u64)q.z | ((u64)q.w << 32);
                g = le; r8 = 0;
            } else {
                edge = (__shfl_sync(0xffffffffu, m, g) & ~7ull) + r8;
                zwg = __shfl_sync(0xffffffffu, zw, g);
            }
            const u32 t = t0 + le * world;
            u32 adv = 32;
            if (dupm) {
                // Two groups of this chunk drew the same line (tables of a few hundred lines only). Their events must not
                // share a dense step — the later one has to see the weight the earlier one wrote — so the step is cut
                // in front of the first lane whose line already appears in it under another group; the rest is redone.
                const unsigned cm0 = __ballot_sync(0xffffffffu, cand);
                unsigned pe = 1u << lane;
                if (cand) pe = __match_any_sync(cm0, IID ? (u32)edge : (u32)(edge >> LOGB));
                const u32 g_first = __shfl_sync(0xffffffffu, g, __ffs(pe) - 1);
                const unsigned cut = __ballot_sync(0xffffffffu, cand && g != g_first);
                if (cut) { adv = __ffs(cut) - 1; cand = cand && lane < adv; }
            }
            float w = __uint_as_float(sy.z);
            bool skip = true, want = false;
            int gap = 0, fv = (int)sy.w;
            if (cand) {
                if (IID ? dupm != 0u : ((dupm >> g) & 1u) != 0u) w = __ldcg(&d.syn[edge].w);
                if (spilled) fv = __ldcg(d.fire32 + sy.y);                                  // brain.metal:79
                gap = (int)t - fv;
                skip = (u32)(gap < 0 ? -gap : gap) <= refr;                                 // brain.metal:79-83
                if (!spilled)
                    for (u32 f = 0; f < nf; ++f) skip = skip || fl_dst[f] == sy.y;          // fired earlier in this chunk: refractory
                const float u = kp.release_rng == ABNN_RNG_XORSHIFT ? rand01_xorshift((ev0 + le) ^ (u32)(clock + t))
                                                                    : u01_24(release_word((u32)zwg, SB, (g % LPG) * B + r8));
                want = !skip && release_test(kp, w, u);                                     // brain.metal:91-92
            }
            // in-warp order: the first event of a destination that fires blocks its later events (see header). An event is
            // blocked iff an EARLIER lane of the step wants to fire on the same destination (the earliest such lane does
            // fire); one shuffle per wanting lane (1-2 per step) instead of a match.any over up to 32 distinct destinations.
            bool blocked = false;
            for (unsigned wm = __ballot_sync(0xffffffffu, want); wm; wm &= wm - 1) {
                const int wl = __ffs(wm) - 1;
                const u32 dw = __shfl_sync(0xffffffffu, sy.y, wl);
                blocked = blocked || (dw == sy.y && (int)lane > wl);
            }
            const bool fired = want && !blocked, gated = cand && !skip && !blocked;
            if (gated) {
                float isi = (float)(u32)(gap < 0 ? -gap : gap);                             // brain.metal:116 (now - ld)
                if (fv == FIRE32_ANCIENT || fv == FIRE32_FUTURE) {   // beyond 2^30 ticks: the exact inter-spike interval
                    const u64 ld = __ldcg(d.live + sy.y), now = clock + t;
                    isi = (float)(ld <= now ? now - ld : ld - now);
                }
                const float w_new = plasticity_f(kp, w, fired, R, rbar, isi);               // brain.metal:101-121
                __stcg(&d.syn[edge].w, w_new);                                              // brain.metal:122
                stage_prune(kp, d, edge, w_new);
            }
            if (fired) {
                atomicMax(d.fire32 + sy.y, (int)t);                                         // brain.metal:125-126
                if (GROW) stage_growth(kp, d, event_base + ev0 + le, tick_base + t, sy.x, trial_word((u32)(zwg >> 32), SB, (g % LPG) * B + r8));
            }
            n_gated += gated; n_fired += fired;
            const unsigned fm = __ballot_sync(0xffffffffu, fired);
            if (fm) {
                const u32 at = nf + __popc(fm & lt);
                if (fired && at < LINE_FIRE_CAP) fl_dst[at] = sy.y;
                nf += __popc(fm);
                if (nf > LINE_FIRE_CAP) { nf = LINE_FIRE_CAP; spilled = true; }
                __syncwarp();
            }
            if (dupm | (u32)spilled) __threadfence();        // rare: make this step's writes visible to the re-reads
            j += adv;
            if (adv == 32) { cand = cand_n; le = le_n; sy = sy_n; }
            else {                                           // a cut step: take up again behind it
                cand = j + lane < nC;
                if (cand) {
                    le = queue[j + lane];
                    sy = in_cq ? cq[j + lane] : *reinterpret_cast<const uint4*>(stage + le * 16);
                }
            }
        }
        __syncwarp();                                        // every lane is done with the stage and the queues
        if (!early) { okm_next = stage_issue(m_next, c_next); dupm_next = dup_mask(m_next); }
        c = c_next; m = m_next; zw = zw_next; okm = okm_next; dupm = dupm_next;
    }
    flush_counters(d, n_cand, n_gated, n_fired, s_cnt);
}

// Before the pass: gate words of neurons [s0, s1) from the snapshot `src` (as k_build_slack), fire32 / vis32 of the owned
// neurons [o0, o1) from the 64-bit lastFired.
__global__ void __launch_bounds__(256) k_prepare32(const __grid_constant__ KParams kp, const DevPtrs d, const u64* src, u64 s0, u64 s1,
                                                   u64 o0, u64 o1)
{
    const u64 clock = d.sc->clock;
    const u64 step = (u64)gridDim.x * blockDim.x, tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (u64 n = s0 + tid; n < s1; n += step) d.slack[n] = slack_word(clock, src[n], kp.window_pre);
    for (u64 n = o0 + tid; n < o1; n += step) { d.fire32[n] = fire_word(clock, d.live[n]); d.vis32[n] = 0u; }
}
// After the pass (k_end_pass has advanced the clock), one sweep over the owned neurons [o0, o1): the pass's fires and
// visits go back into the 64-bit arrays, and the three 32-bit words of the NEXT pass are written in the same breath
// (gate word as k_build_slack would build it from the new lastFired, fire word relative to the new clock, visit word
// cleared) — so a steady run never reads the 64-bit arrays at the start of a pass.
// A non-negative fire32 is lastFired - pass start: a fire of this pass, or an uploaded future timestamp that no fire
// overtook (the value the 64-bit array already holds). Single GPU: snapshot == lastFired between passes, kept here.
__global__ void __launch_bounds__(256) k_fold_prepare32(const __grid_constant__ KParams kp, const DevPtrs d, u64 o0, u64 o1)
{
    const u64 clock = d.sc->clock, start = clock - d.sc->last_pass_ticks;
    for (u64 n = o0 + (u64)blockIdx.x * blockDim.x + threadIdx.x; n < o1; n += (u64)gridDim.x * blockDim.x) {
        const int f = d.fire32[n];
        const u32 v = d.vis32[n];
        u64 lf;
        if (f == FIRE32_ANCIENT || f == FIRE32_FUTURE) lf = d.live[n];            // beyond 2^30 ticks: the 64-bit value is the truth
        else {
            lf = start + (u64)(long long)f;
            if (f >= 0) { d.live[n] = lf; d.view[n] = lf; }                          // brain.metal:125-126, folded
        }
        if (v) { const u64 tv = start + v - 1u; if (d.visited[n] < tv) d.visited[n] = tv; }   // README.md:84, folded
        d.slack[n] = slack_word(clock, lf, kp.window_pre);
        d.fire32[n] = fire_word(clock, lf);
        d.vis32[n] = 0u;
    }
}
cudaError_t launch_prepare32(const KParams& kp, const DevPtrs& d, const u64* src, u64 s0, u64 s1, u64 o0, u64 o1, cudaStream_t st)
{
    if (s1 < s0) s1 = s0;
    if (o1 < o0) o1 = o0;
    const u64 span = s1 - s0 > o1 - o0 ? s1 - s0 : o1 - o0;
    if (!span) return cudaSuccess;
    u64 blocks = (span + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_prepare32<<<(unsigned)blocks, 256, 0, st>>>(kp, d, src, s0, s1, o0, o1);
    return cudaGetLastError();
}
cudaError_t launch_fold_prepare32(const KParams& kp, const DevPtrs& d, u64 o0, u64 o1, cudaStream_t st)
{
    if (o1 <= o0) return cudaSuccess;
    u64 blocks = (o1 - o0 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_fold_prepare32<<<(unsigned)blocks, 256, 0, st>>>(kp, d, o0, o1);
    return cudaGetLastError();
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
    sc->gated = 0; sc->fired = 0; sc->cands = 0; sc->grown_pass = 0; sc->fires_claimed = 0; sc->chunk_ticket = 0;
}

// ---- launchers -----------------------------------------------------------------------------------
template <int SAMPLER, int VISITS, int U>
static cudaError_t launch_parallel_u(const KParams& kp, const DevPtrs& d, int sm_count, int per_sm_req, cudaStream_t st)
{
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_traverse_parallel<SAMPLER, VISITS, U>, 256, 0);
    if (per_sm < 1) per_sm = 1;
    if (per_sm_req > 0 && per_sm_req < per_sm) per_sm = per_sm_req;
    const u64 tiles = (kp.count + 256ull * U - 1) / (256ull * U);   // 8 warps x 32*U events
    u64 grid = (u64)sm_count * per_sm;                 // persistent: a whole number of waves
    if (grid > tiles) grid = tiles;
    // Unordered execution only approximates event order at the granularity of what is in flight; keep
    // that window below 1/16 of the pass so small passes stay close to the ordered semantics.
    if (grid > tiles / 16 + 1) grid = tiles / 16 + 1;
    if (grid == 0) return cudaSuccess;
    k_traverse_parallel<SAMPLER, VISITS, U><<<(unsigned)grid, 256, 0, st>>>(kp, d);
    return cudaGetLastError();
}
template <int SAMPLER, int VISITS>
static cudaError_t launch_parallel_t(const KParams& kp, const DevPtrs& d, int sm_count, cudaStream_t st)
{
    // tuning knobs (read once): events in flight per thread, CTAs per SM
    static const int u = tune_env("ABNN_TRAV_U") ? atoi(tune_env("ABNN_TRAV_U")) : 4;
    static const int bps = tune_env("ABNN_TRAV_CTAS") ? atoi(tune_env("ABNN_TRAV_CTAS")) : 0;
    switch (u) {
        case 1:  return launch_parallel_u<SAMPLER, VISITS, 1>(kp, d, sm_count, bps, st);
        case 2:  return launch_parallel_u<SAMPLER, VISITS, 2>(kp, d, sm_count, bps, st);
        case 8:  return launch_parallel_u<SAMPLER, VISITS, 8>(kp, d, sm_count, bps, st);
        default: return launch_parallel_u<SAMPLER, VISITS, 4>(kp, d, sm_count, bps, st);
    }
}


template <int VISITS, int SLACK>
static cudaError_t launch_line(const KParams& kp, const DevPtrs& d, int sm_count, cudaStream_t st)
{
    static const int bps = tune_env("ABNN_TRAV_CTAS") ? atoi(tune_env("ABNN_TRAV_CTAS")) : 0;
    // function attributes are per device: one process may own handles on several GPUs
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_traverse_line<VISITS, SLACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LINE_SMEM);
        if (e != cudaSuccess) return e;
        if (tune_env("ABNN_LINE_CARVEOUT"))      // measurements only: shared-memory share of the L1/shared array, percent
            cudaFuncSetAttribute(k_traverse_line<VISITS, SLACK>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(tune_env("ABNN_LINE_CARVEOUT")));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    if (!kp.count || !kp.n_local) return cudaSuccess;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_traverse_line<VISITS, SLACK>, LINE_WARPS * 32, LINE_SMEM);
    if (per_sm < 1) per_sm = 1;
    if (bps > 0 && bps < per_sm) per_sm = bps;
    const u64 chunks = (kp.count + 255) / 256;
    u64 grid = (u64)sm_count * per_sm;
    if (grid > (chunks + LINE_WARPS - 1) / LINE_WARPS) grid = (chunks + LINE_WARPS - 1) / LINE_WARPS;
    // Events in flight at once (grid x 8 warps x 256) execute unordered. The only cross-warp order the semantics
    // depend on is the refractory gate (a fire blocks the later events of its neuron for `refractory` ticks), and
    // only when the refractory period is longer than a chunk (inside a chunk the warp orders exactly): keep the
    // in-flight window below a quarter of it. Measured at the 10k-neuron toy shape: fired count 3.7 % above the
    // oracle with 63k ticks in flight against a 100k-tick refractory period, 1.0 % with 16k (profiles/r1_notes.md).
    const u64 refractory_chunks = kp.refractory / (256ull * (kp.world ? kp.world : 1));
    if (kp.clock_mode != ABNN_CLOCK_PER_PASS && refractory_chunks >= 1) {
        u64 lim = refractory_chunks / (4 * LINE_WARPS);
        if (lim < 1) lim = 1;
        if (grid > lim) grid = lim;
    }
    k_traverse_line<VISITS, SLACK><<<(unsigned)grid, LINE_WARPS * 32, LINE_SMEM, st>>>(kp, d);
    return cudaGetLastError();
}

template <int VISITS, int GROW, int SB>
static cudaError_t launch_line32(const KParams& kp, const DevPtrs& d, int sm_count, cudaStream_t st)
{
    static bool configured[64] = {};                         // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_traverse_line32<VISITS, GROW, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LINE32_SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    if (!kp.count || !kp.n_local) return cudaSuccess;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_traverse_line32<VISITS, GROW, SB>, LINE_WARPS * 32, LINE32_SMEM);
    if (per_sm < 1) per_sm = 1;
    const u64 chunks = (kp.count + LINE32_EPC - 1) / LINE32_EPC;
    u64 grid = (u64)sm_count * per_sm;
    if (grid > (chunks + LINE_WARPS - 1) / LINE_WARPS) grid = (chunks + LINE_WARPS - 1) / LINE_WARPS;
    // Events in flight at once execute unordered; the one cross-warp order the semantics depend on is the refractory gate,
    // and this kernel tests it against the fire word read at the START of a chunk: keep the in-flight window below 1/16 of
    // the refractory period (measured at the 10k-neuron toy shape, fired count against the oracle: +4.3 % with a quarter of
    // the period in flight, profiles/r2_notes.md). No limit in practice at the benchmark shape (period = 2 passes).
    u64 lim = kp.refractory / ((u64)LINE32_EPC * kp.world) / (16 * LINE_WARPS);
    if (lim < 1) lim = 1;
    if (grid > lim) grid = lim;
    k_traverse_line32<VISITS, GROW, SB><<<(unsigned)grid, LINE_WARPS * 32, LINE32_SMEM, st>>>(kp, d);
    return cudaGetLastError();
}

template <int LOGB, int VISITS>
static cudaError_t launch_block_t(const KParams& kp, const DevPtrs& d, int sm_count, cudaStream_t st)
{
    static const int bps = tune_env("ABNN_TRAV_CTAS") ? atoi(tune_env("ABNN_TRAV_CTAS")) : 0;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_traverse_block<LOGB, VISITS>, 256, 0);
    if (per_sm < 1) per_sm = 1;
    if (bps > 0 && bps < per_sm) per_sm = bps;
    const u64 chunks = (kp.count + (32ull << LOGB) - 1) / (32ull << LOGB);
    u64 grid = (u64)sm_count * per_sm;
    if (grid > (chunks + 7) / 8) grid = (chunks + 7) / 8;
    if (grid > chunks / (8 * 16) + 1) grid = chunks / (8 * 16) + 1;      // in-flight window <= 1/16 of the pass (see above)
    if (grid == 0) return cudaSuccess;
    k_traverse_block<LOGB, VISITS><<<(unsigned)grid, 256, 0, st>>>(kp, d);
    return cudaGetLastError();
}
template <int VISITS>
static cudaError_t launch_block(const KParams& kp, const DevPtrs& d, int sm_count, cudaStream_t st)
{
    switch (kp.log_block) {
        case 1:  return launch_block_t<1, VISITS>(kp, d, sm_count, st);
        case 2:  return launch_block_t<2, VISITS>(kp, d, sm_count, st);
        case 3:  return launch_block_t<3, VISITS>(kp, d, sm_count, st);
        case 4:  return launch_block_t<4, VISITS>(kp, d, sm_count, st);
        default: return launch_block_t<5, VISITS>(kp, d, sm_count, st);
    }
}

bool line_kernel_selected(const KParams& kp)
{
    static const bool legacy_block = tune_env("ABNN_TRAV_LEGACY_BLOCK") != nullptr;      // A/B measurements only
    return kp.sampler == ABNN_SAMPLER_PHILOX && kp.sample_block == 8 && !legacy_block;
}
// Can this pass run on the 32-bit pass-relative words (slack32 / fire32 / vis32)? Preconditions of the words themselves
// (per-event clock, snapshot src view, no spike budget, ticks and refractory below 2^30; the caller additionally needs the
// gate words: slack_mode) and a kernel that works on them: k_traverse_line32 (sample_block 1 / 8 / 16, refractory >= one
// chunk), the register-staged iid kernel (sample_block 1 otherwise) or the block kernel (everything else except
// sample_block 8, which the 64-bit line kernel serves when the refractory period is shorter than a chunk).
static bool line32_kernel(const KParams& kp)
{
    static const bool legacy_iid = tune_env("ABNN_IID_LEGACY") != nullptr;               // A/B measurements only
    if (kp.sample_block == 1 && legacy_iid) return false;
    return (kp.sample_block == 1 || kp.sample_block == 8 || kp.sample_block == 16) && kp.refractory >= (u64)LINE32_EPC * kp.world;
}
bool line32_selected(const KParams& kp)
{
    const u64 lim = 1ull << 30;
    if (!(kp.sampler == ABNN_SAMPLER_PHILOX && kp.clock_mode == ABNN_CLOCK_PER_EVENT && !kp.budget_on && kp.snapshot &&
          kp.ticks < lim - 1 && kp.refractory < lim)) return false;
    return line32_kernel(kp) || kp.sample_block != 8;
}
cudaError_t launch_traverse_parallel(const KParams& kp, const DevPtrs& d, int sm_count, cudaStream_t st)
{
    const bool ph = kp.sampler == ABNN_SAMPLER_PHILOX, vis = kp.track_visits != 0;
    if (kp.use_line32 && line32_kernel(kp)) {
        const bool grow = kp.p_new > 0.f;
        if (kp.sample_block == 16) {
            if (vis) return grow ? launch_line32<1, 1, 16>(kp, d, sm_count, st) : launch_line32<1, 0, 16>(kp, d, sm_count, st);
            return grow ? launch_line32<0, 1, 16>(kp, d, sm_count, st) : launch_line32<0, 0, 16>(kp, d, sm_count, st);
        }
        if (kp.sample_block == 1) {
            if (vis) return grow ? launch_line32<1, 1, 1>(kp, d, sm_count, st) : launch_line32<1, 0, 1>(kp, d, sm_count, st);
            return grow ? launch_line32<0, 1, 1>(kp, d, sm_count, st) : launch_line32<0, 0, 1>(kp, d, sm_count, st);
        }
        if (vis) return grow ? launch_line32<1, 1, 8>(kp, d, sm_count, st) : launch_line32<1, 0, 8>(kp, d, sm_count, st);
        return grow ? launch_line32<0, 1, 8>(kp, d, sm_count, st) : launch_line32<0, 0, 8>(kp, d, sm_count, st);
    }
    if (line_kernel_selected(kp)) {
        if (kp.use_slack) return vis ? launch_line<1, 1>(kp, d, sm_count, st) : launch_line<0, 1>(kp, d, sm_count, st);
        return vis ? launch_line<1, 0>(kp, d, sm_count, st) : launch_line<0, 0>(kp, d, sm_count, st);
    }
    if (ph && kp.sample_block > 1) return vis ? launch_block<1>(kp, d, sm_count, st) : launch_block<0>(kp, d, sm_count, st);
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
