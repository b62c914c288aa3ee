// oracle/oracle_a.cpp — "Oracle A": the reference's own GPU kernel, compiled verbatim.
//
// TEST INFRASTRUCTURE ONLY. This translation unit #includes the UNMODIFIED reference source
// /root/reference/abnn/src/core/kernels/brain.metal (found through -I at build time, never copied
// into this repo) on top of oracle/metal_shim/metal_stdlib, and sweeps thread ids serially.
// It pins the per-event arithmetic of brain.metal:41-130 (gating, release test, budget, LTP/LTD,
// reward term, homeostasis, clamp, timestamp write) and of renormalise_clock_and_times
// (brain.metal:135-145). Output goes to oracle/_ref/liboracle_a.so (git-ignored).
//
// Two sweep disciplines:
//   hold_clock = 0 : naive serial sweep — thread-group 0 sees now = c, later groups see c+1
//                    (tid 0 ticks the clock on every exit path, brain.metal:75,81,86,129).
//   hold_clock = 1 : every event of pass c sees now = c and the clock becomes c+1 at pass end —
//                    the interleaving a real GPU mostly produces; the product's PER_PASS rule.
#include <cstdint>
#include <cstring>
#include <cstddef>
#include "metal_stdlib"
using namespace metal;
#include "brain.metal"

extern "C" {

struct oracle_a_state {
    SynapsePacked* syn;      // caller-owned, nSyn records
    uint32_t*      lastF;    // caller-owned, nNeuron
    uint32_t*      lastV;    // caller-owned, nNeuron (bound but unused by the kernel)
    uint32_t       clock;
    uint32_t       budget;
    float          reward;
    float          rbar;
};

// One pass = Brain::encode_traversal (brain.cpp:87-122): budget reset, then `threads` kernel threads
// rounded up to the 256-wide thread-group like dispatchThreads at brain.cpp:116-118.
void oracle_a_pass(oracle_a_state* st, uint32_t nSyn, uint32_t events, uint32_t maxSpikes,
                   float aLTP, float aLTD, float wMin, float wMax, int hold_clock)
{
    const uint32_t tg = 256;
    const uint32_t threads = ((events + tg - 1) / tg) * tg;
    const uint32_t tauVis = 50000, tauPre = 50000;          // brain.cpp:102 (bound, unused)
    st->budget = maxSpikes;                                  // brain.cpp:90
    const uint32_t c0 = st->clock;
    for (uint32_t t = 0; t < threads; ++t) {
        monte_carlo_traversal(st->syn, st->lastF, st->lastV, &st->clock, nSyn, tauVis, tauPre,
                              aLTP, aLTD, wMin, wMax, &st->budget, &st->reward, &st->rbar,
                              uint3{t % tg, 0, 0}, uint3{t, 0, 0}, uint3{tg, 1, 1});
        if (hold_clock && t == 0) st->clock = c0;
    }
    if (hold_clock) st->clock = c0 + 1;
}

// Brain::renormalise_if_needed (brain.cpp:125-141) + kernel brain.metal:135-145, serial sweep.
// The kernel re-reads the clock per thread while tid 0 zeroes it: the serial order here makes
// tid 0 subtract `base` and then reset, so later tids would subtract 0. To model the intended
// behaviour (all threads read base before the reset lands) the clock is restored until the end.
void oracle_a_renorm(oracle_a_state* st, uint32_t nNeuron)
{
    const uint32_t base = st->clock;
    const uint32_t tg = 256, threads = ((nNeuron + tg - 1) / tg) * tg;
    for (uint32_t t = 0; t < threads; ++t) {
        renormalise_clock_and_times(st->lastF, st->lastV, &st->clock, nNeuron, t);
        st->clock = base;
    }
    st->clock = 0;
}

uint64_t oracle_a_sizeof_synapse(void) { return sizeof(SynapsePacked); }

}  // extern "C"
