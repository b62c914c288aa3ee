// oracle/oracle_b.cpp — "Oracle B": plain host C++ restatement of the ABNN traversal hot path.
//
// TEST INFRASTRUCTURE ONLY. Nothing under abnn_b200/ may include, link or call this file; it is
// used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as
// the CHECKER and CPU baseline, never as a product path.
//
// Parity status: the per-event arithmetic is pinned against Oracle A (oracle/oracle_a.cpp = the
// reference's brain.metal compiled verbatim) in tests/test_oracle.py; RateFilter / FunctionalDataset
// are pinned against the reference's own sources compiled verbatim (oracle/ref_pieces.cpp) and
// the golden values in tests/golden/. The north-star behaviours the reference does not implement
// (Philox sampling, uint64 per-event clock, lastVisited writes, pruning, synaptogenesis, dst
// sharding) have no reference implementation to pin against: for those this file IS the
// definition ("parity unpinned" for those rows — see DESIGN.md §3).
//
// Every function cites the reference lines it follows (paths relative to /root/reference).
// One ob_handle is one shard (rank of world_size); world_size == 1 is the whole network.
// Everything here is serial and deterministic: events are executed strictly in index order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <random>
#include <thread>
#include <vector>

#include "../include/abnn.h"

namespace {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw — SC'11). Constants as in Random123 / cuRAND
// (/usr/local/cuda/include/curand_philox4x32_x.h:88-91). KATs in tests/test_oracle.py.
inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = 0xD2511F53ull * c0;
        const uint64_t p1 = 0xCD9E8D57ull * c2;
        const uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = uint32_t(p1);
        const uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = uint32_t(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum : uint32_t { STREAM_EVENT = 0, STREAM_GROW = 1, STREAM_INJECT = 2, STREAM_TEACHER = 3, STREAM_INIT = 4 };

inline uint64_t mulhi64(uint64_t a, uint64_t b) { return uint64_t((unsigned __int128)a * b >> 64); }
inline float u01_24(uint32_t x) { return float(x >> 8) * (1.0f / 16777216.0f); }
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }  // metal::clamp

// brain.metal:15-19
inline float rand01_xorshift(uint32_t s)
{
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return float(s & 0xFFFFFF) * (1.0f / 16777216.0f);
}

struct GrowCand { uint64_t order; uint32_t src, dst; };   // order = tick ordinal of the firing event

// Release-draw and synaptogenesis-trial words of the events of one sample group (include/abnn.h, sample_block):
// the group's ONE Philox call q serves all of its events — event `lane` of the group takes
// fmix32(q.z + lane*0x9E3779B9) / fmix32(q.w + lane*0x85EBCA6B) (fmix32 = the MurmurHash3 32-bit finaliser, a
// bijection of the 32-bit word). Groups of one event (iid sampler, SWEEP) use q.z / q.w as they are.
inline uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }
inline uint32_t release_word(const uint32_t q[4], uint64_t B, uint32_t lane) { return B == 1 ? q[2] : fmix32(q[2] + lane * 0x9E3779B9u); }
inline uint32_t trial_word(const uint32_t q[4], uint64_t B, uint32_t lane) { return B == 1 ? q[3] : fmix32(q[3] + lane * 0x85EBCA6Bu); }

}  // namespace

struct ob_handle {
    abnn_params p;
    uint64_t N = 0, slice = 0, lo = 0, hi = 0;
    std::vector<abnn_synapse> syn;          // this shard's records (dst in [lo,hi)), table order
    uint64_t cap = 0;
    std::vector<uint64_t> live;             // lastFired, indexed by global neuron id (owned range is authoritative)
    std::vector<uint64_t> view;             // pass-start snapshot of everyone's lastFired (SNAPSHOT src view)
    std::vector<uint64_t> lastV;            // lastVisited (owned range)
    std::vector<uint64_t> n_local_all;      // live record count of every shard
    uint64_t clock = 0, pass_index = 0, event_base = 0, tick_base = 0, last_pass_ticks = 1;
    float reward = 0.f, rbar = 0.f;
    std::vector<GrowCand> grow;
    uint64_t grow_dropped = 0;
    // structural plasticity with periodic rebuilds (abnn_params.compact_every > 1, include/abnn.h)
    uint64_t struct_steps = 0, n_dead = 0;
    // read-out state (brain-engine.cpp:145-186, rate-filter.h)
    std::vector<float> rate, iir;
    std::vector<std::vector<float>> fir;
    float max_observed = 0.5f;
    uint64_t win_pos = 0, windows_done = 0;
    double last_loss = 0.25;
    bool iir_init = false;
};

namespace {
uint64_t* src_array(ob_handle* h) { return h->p.src_view == ABNN_SRC_SNAPSHOT ? h->view.data() : h->live.data(); }

void recount(ob_handle* h)
{
    if (h->p.world_size == 1) h->n_local_all.assign(1, h->syn.size());
}
bool lazy_mode(const ob_handle* h) { return h->p.compact_every > 1; }
bool rebuild_step(const ob_handle* h) { return !lazy_mode(h) || h->struct_steps % h->p.compact_every == 0; }
void new_table(ob_handle* h) { h->struct_steps = 0; h->n_dead = 0; }

// ABNN_TABLE_DST_SORTED (include/abnn.h): stable sort of the shard's table by dst (counting sort:
// records with equal dst keep their relative order). No reference counterpart — a layout rule of
// the north-star design; table order matters because edge(e) indexes the table.
//
// ABNN_TABLE_DST_INTERLEAVED: the dst-sorted table with the records of every group of 8 consecutive neurons
// (global id >> 3) interleaved — inside a group the order is (rank of the record among its destination's records,
// destination): row r of the group holds the r-th record of each of its neurons that has one. A 128-byte line then
// holds records of 8 adjacent destinations, so no two events of a sample group hit the same neuron.
void sort_table(ob_handle* h)
{
    if (h->p.table_order == ABNN_TABLE_AS_GIVEN || h->syn.empty()) return;
    const uint64_t span = h->hi - h->lo;
    std::vector<uint64_t> start(span + 1, 0);
    for (const auto& s : h->syn) ++start[s.dst - h->lo + 1];
    for (uint64_t d = 0; d < span; ++d) start[d + 1] += start[d];
    std::vector<uint64_t> first(start);                    // start of every destination's run (start[] is advanced below)
    std::vector<abnn_synapse> out(h->syn.size());
    for (const auto& s : h->syn) out[start[s.dst - h->lo]++] = s;
    h->syn.swap(out);
    if (h->p.table_order != ABNN_TABLE_DST_INTERLEAVED) return;
    // neurons per group = records per sample group (8, or 16 for sample_block >= 16): no two events of a group share a neuron
    const uint64_t G = h->p.sample_block >= 16 ? 16 : 8;
    uint64_t o = 0;
    for (uint64_t g0 = h->lo & ~(G - 1); g0 < h->hi; g0 += G) {   // groups are aligned on the GLOBAL neuron id
        uint64_t cnt[16], beg[16], rows = 0;
        for (uint64_t k = 0; k < G; ++k) {
            const uint64_t nid = g0 + k;
            const bool mine = nid >= h->lo && nid < h->hi;
            beg[k] = mine ? first[nid - h->lo] : 0;
            cnt[k] = mine ? first[nid - h->lo + 1] - first[nid - h->lo] : 0;
            rows = std::max(rows, cnt[k]);
        }
        for (uint64_t r = 0; r < rows; ++r)
            for (uint64_t k = 0; k < G; ++k)
                if (r < cnt[k]) out[o++] = h->syn[beg[k] + r];
    }
    h->syn.swap(out);
}
}  // namespace

extern "C" {

void ob_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

int ob_partition(uint64_t n_neuron, uint32_t world, uint32_t rank, uint64_t* lo, uint64_t* hi)
{
    if (!world || rank >= world) return -1;
    const uint64_t slice = (n_neuron + world - 1) / world;
    *lo = std::min<uint64_t>(n_neuron, slice * rank);
    *hi = std::min<uint64_t>(n_neuron, slice * (rank + 1));
    return 0;
}

int ob_event_share(uint64_t events, uint64_t n_global, uint64_t before, uint64_t n_local,
                   uint64_t* first, uint64_t* count)
{
    if (n_global == 0) { *first = 0; *count = 0; return 0; }
    const uint64_t a = uint64_t((unsigned __int128)events * before / n_global);
    const uint64_t b = uint64_t((unsigned __int128)events * (before + n_local) / n_global);
    *first = a; *count = b - a;
    return 0;
}

// Brain::Brain + build_buffers (brain.cpp:21-27,52-69): zero-filled state, reward = rBar = 0, clock = 0.
ob_handle* ob_create(const abnn_params* p)
{
    if (!p || p->struct_size != sizeof(abnn_params) || p->world_size == 0 || p->rank >= p->world_size) return nullptr;
    ob_handle* h = new ob_handle;
    h->p = *p;
    h->N = uint64_t(p->n_input) + p->n_output + p->n_hidden;
    h->slice = (h->N + p->world_size - 1) / p->world_size;
    ob_partition(h->N, p->world_size, p->rank, &h->lo, &h->hi);
    h->live.assign(h->N, 0); h->view.assign(h->N, 0); h->lastV.assign(h->N, 0);
    h->n_local_all.assign(p->world_size, 0);
    // capacity rule of the C-ABI (include/abnn.h: syn_capacity 0 = this rank's share of n_syn, plus slack when sharded)
    h->cap = p->syn_capacity;
    if (!h->cap)
        h->cap = p->world_size == 1 ? p->n_syn
                                    : (p->n_syn + p->world_size - 1) / p->world_size + p->n_syn / (16ull * p->world_size) +
                                          65536 + uint64_t(p->n_input) * p->n_output;
    if (!h->cap) h->cap = 1;
    h->max_observed = p->peak_init;
    h->last_loss = p->loss0;
    h->rate.assign(p->n_output, 0.f);
    return h;
}
void ob_destroy(ob_handle* h) { delete h; }

// Every shard must know every shard's record count (event shares, pass length).
int ob_set_shard_counts(ob_handle* h, const uint64_t* n_local_all)
{
    h->n_local_all.assign(n_local_all, n_local_all + h->p.world_size);
    return 0;
}
uint64_t ob_n_syn_local(ob_handle* h) { return h->syn.size(); }

// Keep, in table order, the records whose dst this shard owns.
int ob_upload_synapses(ob_handle* h, const abnn_synapse* s, uint64_t n)
{
    h->syn.clear();
    for (uint64_t i = 0; i < n; ++i)
        if (s[i].dst >= h->lo && s[i].dst < h->hi) h->syn.push_back(s[i]);
    if (h->syn.size() > h->cap) return ABNN_ERR_CAPACITY;
    sort_table(h);
    recount(h);
    new_table(h);
    return 0;
}
int ob_download_synapses(ob_handle* h, abnn_synapse* out, uint64_t cap, uint64_t* n_out)
{
    *n_out = h->syn.size();
    if (cap < h->syn.size()) return ABNN_ERR_CAPACITY;
    std::memcpy(out, h->syn.data(), h->syn.size() * sizeof(abnn_synapse));
    return 0;
}

// build_random_graph (brain-engine.cpp:31-53), restated: mt19937(seed) [reference: seed 1],
// wIn ~ U[.4,.8), wHH ~ U[.1,.2); dense input->output first, then hidden->hidden with the draw
// order hid(src), hid(dst), wHH (braced-init-list evaluation is left to right).
// Uses this toolchain's libstdc++ distributions, like the reference uses its platform's.
static void build_reference_graph(std::vector<abnn_synapse>& out, uint32_t n_in, uint32_t n_out,
                                  uint64_t n_neuron, uint64_t n_syn, uint64_t seed)
{
    std::mt19937 gen((uint32_t)seed);
    std::uniform_real_distribution<float> wIn(0.4f, 0.8f), wHH(0.1f, 0.2f);
    out.resize(n_syn);
    uint64_t idx = 0;
    for (uint32_t i = 0; i < n_in && idx < n_syn; ++i)
        for (uint32_t o = 0; o < n_out && idx < n_syn; ++o)
            out[idx++] = abnn_synapse{i, n_in + o, wIn(gen), 0.f};
    std::uniform_int_distribution<uint32_t> hid(n_in + n_out, uint32_t(n_neuron - 1));
    while (idx < n_syn) {
        const uint32_t a = hid(gen);
        const uint32_t b = hid(gen);
        const float    w = wHH(gen);
        out[idx++] = abnn_synapse{a, b, w, 0.f};
    }
}

// README.md:134-135 "Erdős–Rényi ... weights ~ Beta(2,8)". Edge g (global index) is a pure
// function of (seed, g): 4 Philox calls -> src, dst uniform; w = 2nd smallest of 9 uniforms
// (the a-th order statistic of a+b-1 uniforms is Beta(a,b); exact, no transcendental functions).
// Sharded: shard k generates global edges [floor(n*k/G), floor(n*(k+1)/G)) with dst uniform over
// its own neuron range.
static abnn_synapse er_beta_edge(uint64_t g, uint64_t seed, uint64_t N, uint64_t dlo, uint64_t dhi)
{
    uint32_t r[16];
    for (uint32_t c = 0; c < 4; ++c)
        philox4x32_10(uint32_t(g), uint32_t(g >> 32), c, STREAM_INIT, uint32_t(seed), uint32_t(seed >> 32), r + 4 * c);
    abnn_synapse s;
    s.src = uint32_t(mulhi64((uint64_t(r[0]) << 32) | r[1], N));
    s.dst = uint32_t(dlo + mulhi64((uint64_t(r[2]) << 32) | r[3], dhi - dlo));
    float m1 = 2.f, m2 = 2.f;                       // two smallest of 9
    for (int j = 0; j < 9; ++j) {
        const float u = u01_24(r[4 + j]);
        if (u < m1) { m2 = m1; m1 = u; } else if (u < m2) { m2 = u; }
    }
    s.w = m2; s.pad = 0.f;
    return s;
}

int ob_init_graph(ob_handle* h, uint32_t kind, uint64_t seed)
{
    const abnn_params& p = h->p;
    if (kind == ABNN_GRAPH_REFERENCE) {
        std::vector<abnn_synapse> all;
        build_reference_graph(all, p.n_input, p.n_output, h->N, p.n_syn, seed);
        int rc = ob_upload_synapses(h, all.data(), all.size());
        return rc;
    }
    if (kind == ABNN_GRAPH_ER_BETA) {
        const uint64_t g0 = uint64_t((unsigned __int128)p.n_syn * p.rank / p.world_size);
        const uint64_t g1 = uint64_t((unsigned __int128)p.n_syn * (p.rank + 1) / p.world_size);
        h->syn.resize(g1 - g0);
        for (uint64_t g = g0; g < g1; ++g) h->syn[g - g0] = er_beta_edge(g, seed, h->N, h->lo, h->hi);
        if (p.world_size > 1)
            for (uint32_t k = 0; k < p.world_size; ++k)
                h->n_local_all[k] = uint64_t((unsigned __int128)p.n_syn * (k + 1) / p.world_size) -
                                    uint64_t((unsigned __int128)p.n_syn * k / p.world_size);
        sort_table(h);
        recount(h);
        new_table(h);
        return 0;
    }
    return ABNN_ERR_INVALID;
}

// Brain::inject_inputs (brain.cpp:73-83). pTick keeps the reference's float arithmetic
// (hz * kTickNS(uint32 1000) * NSEC_PER_SEC(1000000000ull) = 1e15 for hz = 1000); the host
// mt19937 draw is replaced by Philox(seed; pass, i).
int ob_inject_inputs(ob_handle* h, const float* v, uint32_t n, float hz)
{
    if (n != h->p.n_input) return ABNN_ERR_INVALID;
    const float pTick = hz * 1000u * 1000000000ull;
    const uint64_t now = h->clock;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t r[4];
        philox4x32_10(uint32_t(h->pass_index), uint32_t(h->pass_index >> 32), i, STREAM_INJECT,
                      uint32_t(h->p.seed), uint32_t(h->p.seed >> 32), r);
        if (u01_24(r[0]) < pTick * v[i]) { h->live[i] = now; h->view[i] = now; }
    }
    return 0;
}

// Teacher forcing (brain-engine.cpp:119-134): p = expected[o]*rate; spike iff u < p and the
// output is more than teacher_gap ticks past its last spike.
int ob_teacher_force(ob_handle* h, const float* expected, uint32_t n, float rate)
{
    if (n != h->p.n_output) return ABNN_ERR_INVALID;
    const uint64_t now = h->clock;
    const uint64_t* lf_read = src_array(h);
    for (uint32_t o = 0; o < n; ++o) {
        uint32_t r[4];
        philox4x32_10(uint32_t(h->pass_index), uint32_t(h->pass_index >> 32), o, STREAM_TEACHER,
                      uint32_t(h->p.seed), uint32_t(h->p.seed >> 32), r);
        const float pr = expected[o] * rate;
        const uint64_t id = h->p.n_input + o;
        if (u01_24(r[0]) < pr && (now - lf_read[id] > h->p.teacher_gap)) { h->live[id] = now; h->view[id] = now; }
    }
    return 0;
}

int ob_set_reward(ob_handle* h, float r) { h->reward = r; return 0; }
int ob_get_reward(ob_handle* h, float* r, float* rbar) { if (r) *r = h->reward; if (rbar) *rbar = h->rbar; return 0; }

// One pass, events executed strictly in order. Per-event body = brain.metal:70-126 with the
// build decisions of SURVEY.md §8.0 (runtime parameters, uint64 timestamps, saturating budget,
// lastVisited write from README.md:84, Philox sampling from README.md:77).
int ob_run_pass(ob_handle* h, uint64_t events, abnn_pass_stats* st)
{
    const abnn_params& p = h->p;
    const uint32_t G = p.world_size, k = p.rank;
    uint64_t n_global = 0, before = 0, max_count = 0;
    for (uint32_t j = 0; j < G; ++j) { if (j < k) before += h->n_local_all[j]; n_global += h->n_local_all[j]; }
    uint64_t first = 0, count = 0;
    {
        uint64_t acc = 0;
        for (uint32_t j = 0; j < G; ++j) {
            uint64_t f, c; ob_event_share(events, n_global, acc, h->n_local_all[j], &f, &c);
            if (j == k) { first = f; count = c; }
            max_count = std::max(max_count, c);
            acc += h->n_local_all[j];
        }
    }
    (void)first;
    const uint64_t n_local = h->syn.size();
    const uint64_t B = p.sample_block ? p.sample_block : 1;
    const bool budget_on = p.max_spikes_per_pass != 0;
    uint64_t fires_left = budget_on
        ? (uint64_t(p.max_spikes_per_pass) * (k + 1) / G - uint64_t(p.max_spikes_per_pass) * k / G) : 0;   // brain.cpp:90
    const uint64_t* srcv = src_array(h);
    uint64_t gated = 0, fired_n = 0, cands = 0, grown = 0;
    const float R = h->reward;                                   // brain.metal:105

    for (uint64_t i = 0; i < count; ++i) {
        // K1: event -> synapse. One Philox call per sample group (B events): .x.y -> position, .z/.w -> the release
        // and growth words of the group's events (release_word / trial_word above).
        const uint64_t eid = h->event_base + i;
        const uint64_t Bw = p.sampler == ABNN_SAMPLER_PHILOX ? B : 1;        // events per Philox call
        const uint32_t lane = uint32_t(i % Bw);
        const uint64_t eid0 = eid - lane;
        uint32_t q[4] = {0, 0, 0, 0};
        const bool need_philox = p.sampler == ABNN_SAMPLER_PHILOX || p.release_rng == ABNN_RNG_PHILOX || p.p_new > 0.f;
        if (need_philox)
            philox4x32_10(uint32_t(eid0), uint32_t(eid0 >> 32), k, STREAM_EVENT, uint32_t(p.seed), uint32_t(p.seed >> 32), q);
        uint64_t edge;
        if (p.sampler == ABNN_SAMPLER_SWEEP) { edge = i; if (edge >= n_local) continue; }   // brain.metal:60-61
        else if (B == 1) { if (!n_local) break; edge = mulhi64((uint64_t(q[0]) << 32) | q[1], n_local); }   // README.md:77
        else {
            // block sampler (include/abnn.h sample_block): the group's first event draws the block
            if (!n_local) break;
            edge = B * mulhi64((uint64_t(q[0]) << 32) | q[1], (n_local + B - 1) / B) + lane;
            if (edge >= n_local) continue;
        }
        // K2: clock
        const uint64_t now = p.clock_mode == ABNN_CLOCK_PER_PASS ? h->clock : h->clock + i * G + k;
        abnn_synapse s = h->syn[edge];                           // brain.metal:70
        if (s.src == ABNN_DEAD_SRC) continue;                    // pruned, waiting for the next rebuild: the event does nothing
        if (p.track_visits && h->lastV[s.dst] < now) h->lastV[s.dst] = now;   // README.md:84
        // K3: gating
        const uint64_t lp = srcv[s.src];
        if (now - lp > p.window_pre) continue;                   // brain.metal:73-77
        ++cands;
        const uint64_t ld = h->live[s.dst];
        if (now - ld <= p.refractory) continue;                  // brain.metal:79-83
        if (budget_on && fires_left == 0) continue;              // brain.metal:85-88
        // K5: release
        const float pr = clampf(s.w * s.w * p.base_scale, 0.f, 1.f);
        const float u = p.release_rng == ABNN_RNG_XORSHIFT ? rand01_xorshift(uint32_t(i) ^ uint32_t(now)) : u01_24(release_word(q, Bw, lane));
        bool fired = pr > u;                                     // brain.metal:91-92
        if (fired && budget_on) --fires_left;                    // brain.metal:95-98 (serial: never loses the race)
        // K6: plasticity
        float dW = fired ? (p.a_ltp * (1.f - s.w)) : (-p.a_ltd * s.w);      // brain.metal:101-102
        const float rBar = h->rbar;
        dW += p.eta_reward * (R - rBar) * (fired ? 1.0f : 0.0f);            // brain.metal:105-107
        if (p.rbar_mode == ABNN_RBAR_METAL_TID0 && i == 0 && k == 0)
            h->rbar = rBar + p.alpha_rbar * (R - rBar);                     // brain.metal:110-113
        const float isi = float(now - ld);
        const float estHz = isi > 0.f ? p.home_tick_hz / isi : 0.f;         // brain.metal:116-117
        dW += p.eta_home * (p.target_rate_hz - estHz) * s.w;                // brain.metal:118
        s.w = clampf(s.w + dW, p.w_min, p.w_max);                           // brain.metal:121
        h->syn[edge].w = s.w;                                               // brain.metal:122
        ++gated;
        if (fired) {
            if (h->live[s.dst] < now) h->live[s.dst] = now;                 // brain.metal:125-126 (max: README.md:106 order-free)
            ++fired_n;
            // README.md:125 synaptogenesis: "rand() < p_new on fire -> append (src, dst')"
            if (p.p_new > 0.f && float(trial_word(q, Bw, lane)) * (1.0f / 4294967296.0f) < p.p_new) {
                uint32_t g[4];
                philox4x32_10(uint32_t(eid), uint32_t(eid >> 32), k, STREAM_GROW, uint32_t(p.seed), uint32_t(p.seed >> 32), g);
                const uint64_t dsts = h->N - p.n_input;
                const uint32_t nd = uint32_t(p.n_input + mulhi64((uint64_t(g[0]) << 32) | g[1], dsts));
                h->grow.push_back(GrowCand{h->tick_base + i * G + k, s.src, nd});
                ++grown;
            }
        }
    }
    // end of pass
    if (p.rbar_mode == ABNN_RBAR_PASS_STEP) h->rbar = h->rbar + p.alpha_rbar * (R - h->rbar);
    const uint64_t ticks = std::max<uint64_t>(1, uint64_t(G) * max_count);
    if (p.clock_mode == ABNN_CLOCK_PER_PASS) { h->clock += 1; h->last_pass_ticks = 1; }     // brain.metal:129
    else { h->clock += ticks; h->last_pass_ticks = ticks; }                                 // README.md:85
    h->tick_base += ticks;
    h->event_base += max_count;
    h->pass_index += 1;
    if (G == 1 && p.src_view == ABNN_SRC_SNAPSHOT) h->view = h->live;   // single shard: exchange is a copy
    if (st) {
        st->events = count; st->gated = gated; st->fired = fired_n; st->candidates = cands; st->grown = grown;
        st->clock = h->clock; st->device_ms = 0.0;
    }
    return 0;
}

// Multi-shard exchange (the NCCL allgather of SURVEY.md §8e): owned slice out, full view in.
uint64_t* ob_live_ptr(ob_handle* h) { return h->live.data(); }
uint64_t* ob_view_ptr(ob_handle* h) { return h->view.data(); }

// Brain::read_outputs (brain.cpp:145-157): ts != 0 && start <= ts < now, start = now - (ticks of last pass).
int ob_read_outputs(ob_handle* h, uint8_t* spikes, uint32_t n)
{
    if (n != h->p.n_output) return ABNN_ERR_INVALID;
    const uint64_t now = h->clock, span = h->last_pass_ticks;
    const uint64_t start = now > span ? now - span : 0;
    const uint64_t* lf = src_array(h);
    for (uint32_t o = 0; o < n; ++o) {
        const uint64_t ts = lf[h->p.n_input + o];
        spikes[o] = (ts != 0 && ts >= start && ts < now) ? 1 : 0;
    }
    return 0;
}

// brain-engine.cpp:143-186 (rate EMA, RateFilter::process rate-filter.h:22-59, peak normalise,
// windowed loss -> reward).
int ob_readout_filtered(ob_handle* h, const float* expected, float* rates, uint32_t n)
{
    const abnn_params& p = h->p;
    if (n != p.n_output) return ABNN_ERR_INVALID;
    std::vector<uint8_t> sp(n);
    ob_read_outputs(h, sp.data(), n);
    const float alpha = p.rate_alpha;
    for (uint32_t i = 0; i < n; ++i)
        h->rate[i] = (1 - alpha) * h->rate[i] + alpha * (sp[i] ? 1.f : 0.f);          // brain-engine.cpp:149-151
    // RateFilter::process
    if (!h->iir_init) { h->iir = h->rate; h->iir_init = true; }                        // rate-filter.h:24-26
    const double a = p.dt_sec / (p.filter_tau + p.dt_sec);                             // rate-filter.h:29
    for (uint32_t i = 0; i < n; ++i) h->iir[i] += float(a * (h->rate[i] - h->iir[i])); // rate-filter.h:32-34
    std::vector<float> smooth;
    if (p.use_fir) {
        h->fir.push_back(h->iir);                                                      // rate-filter.h:38-41
        if (h->fir.size() > p.fir_size) h->fir.erase(h->fir.begin());
        smooth.assign(n, 0.f);
        for (const auto& fr : h->fir) for (uint32_t i = 0; i < n; ++i) smooth[i] += fr[i];   // rate-filter.h:44-49
        const float inv = 1.0f / float(h->fir.size());
        for (auto& v : smooth) v *= inv;
    } else smooth = h->iir;
    for (float r : smooth) h->max_observed = std::max(h->max_observed, r);             // brain-engine.cpp:156-159
    h->max_observed *= p.peak_decay;
    for (auto& r : smooth) r = std::min(r / h->max_observed, 1.0f);                    // brain-engine.cpp:162-164
    if (expected) {
        ++h->win_pos;                                                                  // brain-engine.cpp:172
        if (h->win_pos == p.reward_window) {
            double loss = 0.0;
            for (uint32_t i = 0; i < n; ++i) { double err = smooth[i] - expected[i]; loss += err * err; }
            loss /= n;
            h->reward = float(h->last_loss - loss);                                    // brain-engine.cpp:180-181
            h->last_loss = loss;
            h->win_pos = 0;
            ++h->windows_done;
        }
    }
    if (rates) std::memcpy(rates, smooth.data(), n * sizeof(float));
    return 0;
}
int ob_get_loss(ob_handle* h, double* l, uint64_t* w) { if (l) *l = h->last_loss; if (w) *w = h->windows_done; return 0; }

// ---- structural plasticity (README.md:120-127) ------------------------------------------------
// Prune: remove records with w < w_prune, order preserved. Returns number removed.
// With compact_every = K > 1 (README.md:122-124 "remove, compact periodically") only the structural steps 0, K, 2K, ...
// since the table was made remove records; in between a pruned record is marked dead in place (src = ABNN_DEAD_SRC, its
// slot is still sampled, the event does nothing). Returns the number of records that were alive and are pruned now.
uint64_t ob_prune(ob_handle* h)
{
    const float wp = h->p.w_prune;
    if (!(wp > 0.f)) return 0;
    const size_t n0 = h->syn.size();
    if (!rebuild_step(h)) {
        uint64_t marked = 0;
        for (auto& s : h->syn)
            if (s.src != ABNN_DEAD_SRC && s.w < wp) { s.src = ABNN_DEAD_SRC; ++marked; }
        h->n_dead += marked;
        return marked;
    }
    size_t o = 0;
    for (size_t i = 0; i < n0; ++i) if (!(h->syn[i].w < wp)) h->syn[o++] = h->syn[i];   // dead records carry w < w_prune
    h->syn.resize(o);
    if (o != n0 && h->p.table_order == ABNN_TABLE_DST_INTERLEAVED) sort_table(h);   // the interleaved order is re-derived after every change
    const uint64_t pruned = (n0 - h->n_dead) - o;
    h->n_dead = 0;
    return pruned;
}
// Staged growth candidates of this shard: (order, src, dst) triples, 16 bytes each.
uint64_t ob_grow_count(ob_handle* h) { return h->grow.size(); }
void ob_grow_fetch(ob_handle* h, void* out) { std::memcpy(out, h->grow.data(), h->grow.size() * sizeof(GrowCand)); }
// Append, in `order`, the candidates (from all shards) whose dst this shard owns. Clears the stage.
uint64_t ob_grow_apply(ob_handle* h, const void* cands, uint64_t n, uint64_t* dropped)
{
    std::vector<GrowCand> c((const GrowCand*)cands, (const GrowCand*)cands + n);
    std::stable_sort(c.begin(), c.end(), [](const GrowCand& a, const GrowCand& b) { return a.order < b.order; });
    uint64_t app = 0, drop = 0;
    for (const auto& g : c) {
        if (g.dst < h->lo || g.dst >= h->hi) continue;
        if (h->syn.size() >= h->cap) { ++drop; continue; }
        h->syn.push_back(abnn_synapse{g.src, g.dst, h->p.w_init, 0.f});
        ++app;
    }
    h->grow.clear();
    // eager mode / rebuild step: the new records go to their place in the table order (stable: behind the existing records of
    // their destination, tail records of earlier lazy steps first); lazy step: they stay behind the table, in tick order
    if (rebuild_step(h) ? (app || lazy_mode(h)) : false) sort_table(h);
    h->struct_steps += 1;
    if (dropped) *dropped = drop;
    return app;
}
int ob_prune_and_grow(ob_handle* h, abnn_structural_stats* st)   // single shard convenience
{
    abnn_structural_stats s{};
    s.n_before = h->syn.size();
    s.pruned = ob_prune(h);
    std::vector<GrowCand> c = h->grow;
    s.appended = ob_grow_apply(h, c.data(), c.size(), &s.dropped);
    s.n_after = h->syn.size();
    recount(h);
    if (st) *st = s;
    return 0;
}

int ob_download_timestamps(ob_handle* h, uint64_t* lf, uint64_t* lv)
{
    if (lf) std::memcpy(lf, h->live.data(), h->N * 8);
    if (lv) std::memcpy(lv, h->lastV.data(), h->N * 8);
    return 0;
}
int ob_upload_timestamps(ob_handle* h, const uint64_t* lf, const uint64_t* lv)
{
    if (lf) { std::memcpy(h->live.data(), lf, h->N * 8); std::memcpy(h->view.data(), lf, h->N * 8); }
    if (lv) std::memcpy(h->lastV.data(), lv, h->N * 8);
    return 0;
}
int ob_get_clock(ob_handle* h, uint64_t* c) { *c = h->clock; return 0; }
int ob_set_clock(ob_handle* h, uint64_t c) { h->clock = c; return 0; }

// ---- a world of shards in one process (threaded CPU baseline; mirrors the multi-GPU step) ------
// Runs one pass on every shard concurrently, then performs the timestamp exchange.
int ob_world_run_pass(ob_handle** hs, uint32_t G, uint64_t events, abnn_pass_stats* sum)
{
    std::vector<abnn_pass_stats> st(G);
    std::vector<std::thread> th;
    for (uint32_t k = 0; k < G; ++k) th.emplace_back([&, k] { ob_run_pass(hs[k], events, &st[k]); });
    for (auto& t : th) t.join();
    if (G > 1) {
        // allgather: view[all] <- live[owner's slice]
        for (uint32_t k = 0; k < G; ++k)
            for (uint32_t j = 0; j < G; ++j)
                std::memcpy(hs[k]->view.data() + hs[j]->lo, hs[j]->live.data() + hs[j]->lo, (hs[j]->hi - hs[j]->lo) * 8);
    }
    if (sum) {
        *sum = abnn_pass_stats{};
        for (uint32_t k = 0; k < G; ++k) {
            sum->events += st[k].events; sum->gated += st[k].gated; sum->fired += st[k].fired;
            sum->candidates += st[k].candidates; sum->grown += st[k].grown; sum->clock = st[k].clock;
        }
    }
    return 0;
}
int ob_world_prune_and_grow(ob_handle** hs, uint32_t G, abnn_structural_stats* sum)
{
    std::vector<GrowCand> all;
    abnn_structural_stats s{};
    for (uint32_t k = 0; k < G; ++k) { s.n_before += hs[k]->syn.size(); s.pruned += ob_prune(hs[k]); }
    for (uint32_t k = 0; k < G; ++k) all.insert(all.end(), hs[k]->grow.begin(), hs[k]->grow.end());
    for (uint32_t k = 0; k < G; ++k) { uint64_t d = 0; s.appended += ob_grow_apply(hs[k], all.data(), all.size(), &d); s.dropped += d; }
    std::vector<uint64_t> counts(G);
    for (uint32_t k = 0; k < G; ++k) { counts[k] = hs[k]->syn.size(); s.n_after += counts[k]; }
    for (uint32_t k = 0; k < G; ++k) ob_set_shard_counts(hs[k], counts.data());
    if (sum) *sum = s;
    return 0;
}

// ---- FunctionalDataset (stimulus/functional-dataset.cpp:24-52) with the two lambdas the app
// installs (view-delegate.cpp:37-42): input cos^2(x), expected 0.5*sin(x)+0.5, x passed as float.
struct ob_dataset { uint32_t n_in, n_out; double dt, f, phase, t; };
ob_dataset* ob_dataset_create(uint32_t n_in, uint32_t n_out, double dt, double f)
{
    return new ob_dataset{n_in, n_out, dt, f, 0.0, 0.0};
}
void ob_dataset_destroy(ob_dataset* d) { delete d; }
void ob_dataset_next_input(ob_dataset* d, float* v)
{
    d->phase += d->f * d->dt;                                  // functional-dataset.cpp:29-31
    if (d->phase > 1.0) d->phase -= 1.0;
    d->t += d->dt;
    for (uint32_t i = 0; i < d->n_in; ++i) {
        const double x = double(i) / d->n_in;
        const float a = float(2.0 * M_PI * (x + d->phase));    // functional-dataset.cpp:35
        v[i] = float(cos(a) * cos(a));                         // view-delegate.cpp:37-39 (float arg promotes to double cos)
    }
}
void ob_dataset_next_expected(ob_dataset* d, float* v)
{
    for (uint32_t i = 0; i < d->n_out; ++i) {
        const double x = double(i) / d->n_out;
        const float a = float(2.0 * M_PI * (x + d->phase));    // std::function<float(float)> narrows the double argument
        const float s = float(0.5f * sin(a) + 0.5f);           // view-delegate.cpp:40-42
        v[i] = float(double(s));                               // functional-dataset.cpp:48-49 (double s = f(...); v[i] = s)
    }
}
double ob_dataset_time(ob_dataset* d) { return d->t; }

}  // extern "C"
