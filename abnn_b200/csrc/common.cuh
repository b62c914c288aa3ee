// abnn_b200/csrc/common.cuh — shared definitions of the CUDA library behind include/abnn.h.
// sm_100a only. Built with -fmad=false so the plasticity arithmetic of the rare path is the same
// sequence of IEEE single-precision operations as the reference kernel (brain.metal:91-121).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../include/abnn.h"

namespace abnn {

typedef unsigned long long u64;
typedef unsigned int       u32;

// Measurement knobs (environment variables that switch experimental code paths) exist only in tuning builds
// (nvcc -DABNN_TUNING, e.g. ABNN_NVCC_EXTRA=-DABNN_TUNING python -m abnn_b200.build): the product library reads no
// environment variable — everything that changes its behaviour is a field of abnn_params.
#ifdef ABNN_TUNING
inline const char* tune_env(const char* name) { return getenv(name); }
#else
inline const char* tune_env(const char*) { return nullptr; }
#endif

// Philox counter layout: ctr = (idx.lo, idx.hi, aux, stream), key = seed.
enum : u32 { STREAM_EVENT = 0, STREAM_GROW = 1, STREAM_INJECT = 2, STREAM_TEACHER = 3, STREAM_INIT = 4 };

// Philox4x32-10 (Salmon et al., SC'11). One call serves one event: .x.y -> edge, .z -> release
// draw, .w -> synaptogenesis trial. ~20 IMAD.WIDE + ~30 integer ops per event.
struct Philox4 { u32 x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const u64 p0 = (u64)0xD2511F53u * c0;
        const u64 p1 = (u64)0xCD9E8D57u * c2;
        const u32 n0 = (u32)(p1 >> 32) ^ c1 ^ k0;
        const u32 n2 = (u32)(p0 >> 32) ^ c3 ^ k1;
        c1 = (u32)p1; c3 = (u32)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ u64 mulhi64(u64 a, u64 b)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}
__host__ __device__ __forceinline__ float u01_24(u32 x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
__host__ __device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
// rand01 of the reference kernel (brain.metal:15-19)
__host__ __device__ __forceinline__ float rand01_xorshift(u32 s)
{
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return (float)(s & 0xFFFFFFu) * (1.0f / 16777216.0f);
}

// Release-draw and synaptogenesis-trial words of the events of one sample group (include/abnn.h, sample_block): the
// group's ONE Philox call q serves all of its events — event `lane` of the group takes fmix32(q.z + lane*0x9E3779B9) /
// fmix32(q.w + lane*0x85EBCA6B) (MurmurHash3's 32-bit finaliser, a bijection). Groups of one event (iid sampler, SWEEP)
// use q.z / q.w as they are. Same definition as oracle/oracle_b.cpp:release_word / trial_word.
__host__ __device__ __forceinline__ u32 fmix32(u32 h)
{
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ u32 release_word(u32 qz, u32 group, u32 lane) { return group == 1 ? qz : fmix32(qz + lane * 0x9E3779B9u); }
__host__ __device__ __forceinline__ u32 trial_word(u32 qw, u32 group, u32 lane) { return group == 1 ? qw : fmix32(qw + lane * 0x85EBCA6Bu); }

// 32-bit pre-spike gate word of a neuron for the pass that starts at `clock` (traversal.cu:k_build_slack):
// an event with tick offset t = now - clock passes the window gate (brain.metal:73-77) iff t < word.
// 0 = never in this pass, 0xFFFFFFFE = always, 0xFFFFFFFF = snapshot in the future: take the exact 64-bit test.
constexpr u32 SLACK_EXACT = 0xFFFFFFFFu;
__host__ __device__ __forceinline__ u32 slack_word(u64 clock, u64 last_fired, u64 window_pre)
{
    if (last_fired > clock) return SLACK_EXACT;
    const u64 age = clock - last_fired;
    if (age > window_pre) return 0u;
    const u64 room = window_pre - age;
    return room >= 0xFFFFFFFDull ? 0xFFFFFFFEu : (u32)room + 1u;
}

// 32-bit pass-relative timestamps of the line kernel (traversal.cu:k_traverse_line32). For the pass that starts at
// `clock`, fire32[n] = lastFired[n] - clock as a signed word (<= 0: fired before the pass, >= 0: tick offset of a fire of
// this pass; moved with 32-bit atomicMax) and vis32[n] = 1 + the latest tick offset at which n was visited in this pass
// (0 = not visited). Both are rebuilt from / folded back into the 64-bit arrays around every pass (k_prepare32 /
// k_fold32), so nothing outside the line kernel sees them. Magnitudes are kept below 2^30; the two sentinels mean
// "further away than that": the rare event that needs the exact value reads the 64-bit array.
constexpr int FIRE32_LIMIT = 1 << 30;
constexpr int FIRE32_ANCIENT = -FIRE32_LIMIT;        // lastFired <= clock - 2^30
constexpr int FIRE32_FUTURE = FIRE32_LIMIT - 1;      // lastFired >= clock + 2^30 - 1 (only after an upload of future timestamps)
__host__ __device__ __forceinline__ int fire_word(u64 clock, u64 last_fired)
{
    if (last_fired <= clock) {
        const u64 age = clock - last_fired;
        return age >= (u64)FIRE32_LIMIT ? FIRE32_ANCIENT : -(int)age;
    }
    const u64 ahead = last_fired - clock;
    return ahead >= (u64)FIRE32_FUTURE ? FIRE32_FUTURE : (int)ahead;
}

// Growth candidate staged by a firing event; appended in `order` by the structural step.
struct GrowCand { u64 order; u32 src, dst; };

// Scalars that live on the device so that a pass needs no host round trip
// (the reference keeps them in shared Metal buffers: brain.cpp:54-68).
struct DevScalars {
    u64   clock;             // bufClock_
    u64   pass_index;
    u64   event_base;        // Philox event index of local event 0 of the next pass
    u64   tick_base;         // ordinal of tick 0 of the next pass (growth ordering)
    u64   last_pass_ticks;   // span used by read_outputs
    float reward;            // bufReward_
    float rbar;              // bufRBar_
    // per-pass counters (reset by the pass prologue)
    u64   gated, fired, cands, grown_pass;
    u32   fires_claimed;     // saturating budget (bufBudget_ counts down; this counts up)
    u32   grow_count;        // staged growth candidates since the last structural step
    u32   grow_overflow;
    u32   chunk_ticket;      // next chunk of the pass to hand out (line kernel: dynamic work distribution)
    // read-out state (brain-engine.cpp:145-186; rate-filter.h)
    float max_observed;
    u32   iir_init;
    u32   fir_count, fir_head;
    u32   win_pos;
    u32   pad1;
    u64   windows_done;
    double last_loss;
    // compaction scratch
    u64   compact_total;
    // structural plasticity with periodic rebuilds (abnn_params.compact_every > 1)
    u64   struct_steps;      // structural steps since the table was uploaded / initialised / loaded
    u64   n_sorted;          // records of the ordered region (the rest of the table is the appended tail)
    u64   n_dead;            // dead records waiting for the next rebuild
    u32   prune_count;       // staged prune candidates (records written below w_prune since the last structural step)
    u32   prune_overflow;
};

// Everything a traversal kernel needs, passed by value (constant bank).
struct KParams {
    u64 n_local;         // live records in this rank's table
    u64 count;           // events this rank executes this pass
    u64 ticks;           // ticks this pass spans (G * max_k count_k, >= 1)
    u64 max_count;
    u64 window_pre, refractory;
    u64 n_neuron, neuron_lo, neuron_hi;
    u32 world, rank;
    u32 n_input;
    u32 sampler, release_rng, clock_mode, rbar_mode, track_visits, snapshot;
    u32 budget_on;       // max_spikes_per_pass != 0
    u32 budget_share;    // this rank's share of max_spikes_per_pass
    u32 grow_cap;
    u32 seed_lo, seed_hi;
    u32 sample_block, log_block;   // PHILOX sampler granularity (power of two)
    u64 n_blocks;                  // ceil(n_local / sample_block)
    float base_scale, a_ltp, a_ltd, w_min, w_max, eta_home, target_rate_hz, home_tick_hz, eta_reward, alpha_rbar;
    float p_new;
    u32 use_slack;       // the line kernel reads DevPtrs::slack instead of the 64-bit snapshot
    u32 use_line32;      // the pass runs on the 32-bit pass-relative words fire32 / vis32 (k_traverse_line32, iid and block kernels)
    u32 lazy_prune;      // compact_every > 1: a weight written below w_prune stages its record for the next structural step
    u32 prune_cap;
    float w_prune;
};

struct DevPtrs {
    abnn_synapse* syn;
    u64* view;           // lastFired as seen for src reads (snapshot, or == live)
    u64* live;           // lastFired, authoritative for the owned dst range (indexed by global id)
    u64* visited;        // lastVisited
    u32* slack;          // per-pass 32-bit form of the snapshot for the pre-spike window gate (line kernel), or null
    int* fire32;         // per-pass 32-bit form of lastFired of the owned neurons (k_traverse_line32), or null
    u32* vis32;          // per-pass 32-bit form of lastVisited (k_traverse_line32), or null
    DevScalars* sc;
    GrowCand* grow;
    u64* prune_list;     // staged prune candidates (table indices), or null
};

constexpr u32 DEAD_SRC = ABNN_DEAD_SRC;

// A weight has just been written below w_prune: remember the record for the next structural step (compact_every > 1;
// the step then never has to sweep the table to find what to prune).
__device__ __forceinline__ void stage_prune(const KParams& kp, const DevPtrs& d, u64 edge, float w_new)
{
    if (!kp.lazy_prune || !(w_new < kp.w_prune)) return;
    const u32 slot = atomicAdd(&d.sc->prune_count, 1u);
    if (slot < kp.prune_cap) d.prune_list[slot] = edge;
    else atomicAdd(&d.sc->prune_overflow, 1u);
}

}  // namespace abnn
