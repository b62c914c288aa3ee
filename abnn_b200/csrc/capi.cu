// abnn_b200/csrc/capi.cu — implementation of the C-ABI in include/abnn.h.
// Owns the device state that the reference's `Brain` owns as Metal buffers (brain.cpp:52-69) and
// sequences the kernels of one pass on the handle's stream (Brain::encode_traversal, brain.cpp:87-122).
#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include <nccl.h>

#include "common.cuh"
#include "kernels.h"

using namespace abnn;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(ABNN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)
#define NC(call)                                                                                       \
    do {                                                                                               \
        ncclResult_t r_ = (call);                                                                      \
        if (r_ != ncclSuccess)                                                                         \
            return fail(ABNN_ERR_COMM, std::string(#call) + ": " + ncclGetErrorString(r_));            \
    } while (0)
#define RET(call)                                                                                      \
    do { int rc_ = (call); if (rc_ != 0) return rc_; } while (0)

constexpr int    RING = 16;             // pinned host staging slots for small per-pass vectors
constexpr u64    UPLOAD_CHUNK = 8ull << 20;   // records per upload chunk (128 MB)
constexpr u32    GROW_CAP_DEFAULT = 1u << 20;
constexpr u32    PRUNE_CAP_DEFAULT = 1u << 22;   // staged prune candidates between two structural steps (compact_every > 1)

u32 next_pow2(u32 v) { u32 p = 1; while (p < v) p <<= 1; return p; }

__global__ void k_set_reward(DevScalars* sc, float r) { sc->reward = r; }
__global__ void k_set_clock(DevScalars* sc, u64 c) { sc->clock = c; }
__global__ void k_reset_grow(DevScalars* sc) { sc->grow_count = 0; sc->grow_overflow = 0; }
// bookkeeping of the structural steps (compact_every > 1): step counter, ordered region, dead records, staged candidates
__global__ void k_set_struct(DevScalars* sc, u64 steps, u64 n_sorted, u64 n_dead)
{
    sc->struct_steps = steps; sc->n_sorted = n_sorted; sc->n_dead = n_dead; sc->prune_count = 0; sc->prune_overflow = 0;
}
__global__ void k_pad_grow(GrowCand* c, u32 from, u32 to)
{
    const u32 i = from + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < to) c[i] = GrowCand{~0ull, 0xFFFFFFFFu, 0xFFFFFFFFu};
}

}  // namespace

struct abnn_handle {
    abnn_params p{};
    int device = 0, sm_count = 0;
    size_t l2_bytes = 0, l2_persist = 0;
    u64 N = 0, slice = 0, lo = 0, hi = 0, npad = 0;
    u64 n_local = 0, cap = 0;
    std::vector<u64> n_local_all;
    bool counts_dirty = false;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evj = nullptr, evk = nullptr;
    cudaEvent_t timer[8]{};
    abnn_synapse* d_syn = nullptr;
    abnn_synapse* d_spare = nullptr;      // second table (cap records) kept between sorted growth steps when memory allows
    abnn_synapse* mg_new = nullptr; u32* mg_keys = nullptr; void* mg_tmp = nullptr; u32 mg_cap = 0; size_t mg_tmp_bytes = 0;
    u32* mg_cnt = nullptr; void* mg_scan = nullptr; size_t mg_scan_bytes = 0;   // sorted-growth scratch (merge_grown)
    u32* mg_pruned = nullptr;                                                   // per-neuron removed counts (fused prune + merge)
    size_t mem_total = 0;
    u64* d_ts = nullptr;
    DevPtrs d{};
    u32 grow_cap = 0, grow_buf = 0;
    u64 struct_steps = 0, n_sorted = 0, n_dead = 0;   // host mirror of DevScalars (compact_every > 1)
    GrowCand* d_grow_all = nullptr; u32 grow_all_buf = 0;
    float* d_vec = nullptr;               // RING slots of max(n_input,n_output) floats
    float* h_vec = nullptr;               // pinned mirror
    cudaEvent_t ring_ev[RING]{};
    int ring_pos = 0;
    u32 vec_len = 0;
    ReadoutState rs{};
    abnn_pass_stats* d_stats = nullptr;
    bool l2_user = false;                 // holds a share of the device-wide persisting-L2 set-aside (l2_acquire / l2_release)
    void* h_pin = nullptr;                // pinned scratch (stats, scalars, spikes, rates)
    size_t h_pin_bytes = 0;
    void* d_scratch = nullptr; size_t scratch_bytes = 0;
    abnn_synapse* d_stage = nullptr; u64 stage_cap = 0;
    u64* d_total = nullptr;               // [0]: compaction total, [1..]: misc
    u64* d_counts = nullptr;              // world_size u64 (record counts exchange)
    ncclComm_t comm = nullptr;
    // peer-memory exchange (opt-in, p2p_setup): flag block, IPC mappings of the peers' timestamp allocations and flag blocks
    bool p2p = false;
    u64* d_p2p = nullptr;
    void* peer_ts[P2P_MAX_WORLD]{}; void* peer_flags[P2P_MAX_WORLD]{};
    P2PTable p2p_tab{};
    // EXACT execution scratch (allocated on first use)
    u64* d_xkeys = nullptr; u32* d_xbucket = nullptr; u32* d_xcnt = nullptr; u32* d_xcount = nullptr; void* d_xtmp = nullptr;
    u64 x_cap = 0; size_t x_tmp_bytes = 0;
    bool timing = false;
    // sharded PARALLEL runs exchange the 32-bit slack slices instead of the 64-bit lastFired slices:
    // abnn_engine_step: stimulus frame [in | expected | pTick | rate] and the captured pass
    float* d_frame = nullptr; float* h_frame = nullptr; cudaEvent_t frame_ev[RING]{}; int frame_pos = 0;
    cudaGraphExec_t step_exec = nullptr;
    u64 step_events = 0; std::vector<u64> step_counts; bool step_pre[3]{}, step_post[3]{};
    u64 step_calls = 0, step_replays = 0;
    u64 last_events = 0; std::vector<u64> last_counts; bool last_pre[3]{};   // key of the previous abnn_engine_step call
    bool slack_ready = false;             // d.slack already holds the next pass's gate words (all but the in/out head)
    bool fire_ready = false;              // d.fire32 / d.vis32 of the owned neurons are prepared for the next pass (all but the head)
    bool view_stale = false;              // remote slices of d.view were not refreshed by the last exchange
};

namespace {

int use(abnn_handle* h)
{
    if (!h) return fail(ABNN_ERR_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    return 0;
}

int ensure_scratch(abnn_handle* h, size_t bytes)
{
    if (bytes <= h->scratch_bytes) return 0;
    if (h->d_scratch) CU(cudaFree(h->d_scratch));
    h->d_scratch = nullptr; h->scratch_bytes = 0;
    CU(cudaMalloc(&h->d_scratch, bytes));
    h->scratch_bytes = bytes;
    return 0;
}

// record counts of every rank (event shares, pass length) — exchanged lazily after the table changed
int refresh_counts(abnn_handle* h)
{
    if (h->p.world_size == 1) { h->n_local_all.assign(1, h->n_local); h->counts_dirty = false; return 0; }
    if (!h->counts_dirty) return 0;
    if (!h->comm) return fail(ABNN_ERR_COMM, "world_size > 1 but abnn_comm_init has not been called");
    CU(cudaMemcpyAsync(h->d_counts + h->p.rank, &h->n_local, sizeof(u64), cudaMemcpyHostToDevice, h->st));
    NC(ncclAllGather(h->d_counts + h->p.rank, h->d_counts, 1, ncclUint64, h->comm, h->st));
    CU(cudaMemcpyAsync(h->n_local_all.data(), h->d_counts, sizeof(u64) * h->p.world_size, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    h->counts_dirty = false;
    return 0;
}

void event_share(u64 events, u64 n_global, u64 before, u64 n_local, u64* first, u64* count)
{
    if (!n_global) { *first = 0; *count = 0; return; }
    const u64 a = (u64)((unsigned __int128)events * before / n_global);
    const u64 b = (u64)((unsigned __int128)events * (before + n_local) / n_global);
    *first = a; *count = b - a;
}

KParams make_kparams(const abnn_handle* h, u64 events)
{
    const abnn_params& p = h->p;
    KParams k{};
    k.n_local = h->n_local;
    u64 n_global = 0, acc = 0, maxc = 0, mine = 0;
    for (u64 v : h->n_local_all) n_global += v;
    for (u32 j = 0; j < p.world_size; ++j) {
        u64 f, c;
        event_share(events, n_global, acc, h->n_local_all[j], &f, &c);
        if (j == p.rank) mine = c;
        maxc = std::max(maxc, c);
        acc += h->n_local_all[j];
    }
    k.count = mine; k.max_count = maxc;
    k.ticks = std::max<u64>(1, (u64)p.world_size * maxc);
    k.window_pre = p.window_pre; k.refractory = p.refractory;
    k.n_neuron = h->N; k.neuron_lo = h->lo; k.neuron_hi = h->hi;
    k.world = p.world_size; k.rank = p.rank; k.n_input = p.n_input;
    k.sampler = p.sampler; k.release_rng = p.release_rng; k.clock_mode = p.clock_mode;
    k.rbar_mode = p.rbar_mode; k.track_visits = p.track_visits; k.snapshot = p.src_view == ABNN_SRC_SNAPSHOT;
    k.budget_on = p.max_spikes_per_pass != 0;
    k.budget_share = (u32)((u64)p.max_spikes_per_pass * (p.rank + 1) / p.world_size -
                           (u64)p.max_spikes_per_pass * p.rank / p.world_size);
    k.grow_cap = h->grow_cap;
    k.seed_lo = (u32)p.seed; k.seed_hi = (u32)(p.seed >> 32);
    k.sample_block = p.sample_block ? p.sample_block : 1;
    k.log_block = 0; while ((1u << k.log_block) < k.sample_block) ++k.log_block;
    k.n_blocks = (h->n_local + k.sample_block - 1) / k.sample_block;
    k.base_scale = p.base_scale; k.a_ltp = p.a_ltp; k.a_ltd = p.a_ltd; k.w_min = p.w_min; k.w_max = p.w_max;
    k.eta_home = p.eta_home; k.target_rate_hz = p.target_rate_hz; k.home_tick_hz = p.home_tick_hz;
    k.eta_reward = p.eta_reward; k.alpha_rbar = p.alpha_rbar; k.p_new = p.p_new;
    k.lazy_prune = (p.compact_every > 1 && p.w_prune > 0.f && h->d.prune_list) ? 1u : 0u;
    k.prune_cap = PRUNE_CAP_DEFAULT; k.w_prune = p.w_prune;
    return k;
}

// stage a small host vector into the next ring slot; returns the device pointer of the slot
int stage_vec(abnn_handle* h, const float* v, u32 n, float** d_out)
{
    const int s = h->ring_pos;
    h->ring_pos = (h->ring_pos + 1) % RING;
    CU(cudaEventSynchronize(h->ring_ev[s]));           // slot's previous copy has drained
    float* hp = h->h_vec + (size_t)s * h->vec_len;
    float* dp = h->d_vec + (size_t)s * h->vec_len;
    std::memcpy(hp, v, n * sizeof(float));
    CU(cudaMemcpyAsync(dp, hp, n * sizeof(float), cudaMemcpyHostToDevice, h->st));
    CU(cudaEventRecord(h->ring_ev[s], h->st));
    *d_out = dp;
    return 0;
}

int read_scalars(abnn_handle* h, DevScalars* out)
{
    CU(cudaMemcpyAsync(h->h_pin, h->d.sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    std::memcpy(out, h->h_pin, sizeof(DevScalars));
    return 0;
}

// The line kernel gates on the 32-bit slack words (traversal.cu:k_build_slack) when it can.
bool slack_mode(const abnn_handle* h, const KParams& kp)
{
    static const bool off = tune_env("ABNN_NO_SLACK") != nullptr;
    // the iid and block kernels gate on the 32-bit words too (measured: 16.2 -> 19.2 e9 events/s at the 1B shape, profiles/r2_notes.md)
    const bool kernel_reads_slack = kp.sampler == ABNN_SAMPLER_PHILOX && kp.snapshot;
    return h->p.exec_mode == ABNN_EXEC_PARALLEL && h->d.slack && kernel_reads_slack && kp.ticks < 0xFFFFFFF0ull && !off;
}
// This pass runs on the 32-bit pass-relative words (k_traverse_line32, or the iid / block kernel in their 32-bit form): gate
// words in use and the words' own preconditions (traversal.cu:line32_selected)
bool line32_mode(const abnn_handle* h, const KParams& kp)
{
    static const bool off = tune_env("ABNN_NO_LINE32") != nullptr;
    return slack_mode(h, kp) && h->d.fire32 && line32_selected(kp) && !off;
}

// Peer-memory exchange (exchange.cu): every rank maps the peers' timestamp allocation and flag block through CUDA IPC.
// The 2 x 64-byte handles travel through one ncclAllGather; a rank that cannot map a peer makes EVERY rank fall back to
// the NCCL exchange (ncclAllReduce(min) of the outcome), so the ranks never disagree on which exchange runs.
int p2p_setup(abnn_handle* h)
{
    const u32 W = h->p.world_size, me = h->p.rank;
    if (W < 2 || W > P2P_MAX_WORLD || !h->d.slack) return 0;
    struct Pair { cudaIpcMemHandle_t ts, flags; };
    static_assert(sizeof(Pair) == 128, "two 64-byte IPC handles");
    // A rank-local failure must not keep this rank out of the two collectives below (the peers would wait for it in NCCL
    // forever): it only clears `ok`, which the all-reduce turns into "NCCL exchange on every rank".
    int ok = 1;
    Pair mine{};
    Pair* d_all = nullptr;
    if (cudaMalloc(&h->d_p2p, P2P_WORDS * sizeof(u64)) != cudaSuccess) { h->d_p2p = nullptr; ok = 0; }
    if (ok && (cudaMemsetAsync(h->d_p2p, 0, P2P_WORDS * sizeof(u64), h->st) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine.ts, h->d_ts) != cudaSuccess || cudaIpcGetMemHandle(&mine.flags, h->d_p2p) != cudaSuccess)) ok = 0;
    cudaGetLastError();
    CU(cudaMalloc(&d_all, (size_t)W * sizeof(Pair)));      // 1 KB: if even this fails the handle is unusable anyway
    std::vector<Pair> all(W);
    auto collectives = [&]() -> int {
        CU(cudaMemcpyAsync(d_all + me, &mine, sizeof(Pair), cudaMemcpyHostToDevice, h->st));
        NC(ncclAllGather(d_all + me, d_all, sizeof(Pair), ncclUint8, h->comm, h->st));
        CU(cudaMemcpyAsync(all.data(), d_all, (size_t)W * sizeof(Pair), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        for (u32 r = 0; r < W && ok; ++r) {
            if (r == me) { h->peer_ts[r] = h->d_ts; h->peer_flags[r] = h->d_p2p; continue; }
            if (cudaIpcOpenMemHandle(&h->peer_ts[r], all[r].ts, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&h->peer_flags[r], all[r].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            }
        }
        int* d_ok = reinterpret_cast<int*>(d_all);
        CU(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, h->st));
        NC(ncclAllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, h->comm, h->st));
        CU(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        return 0;
    };
    const int rc = collectives();
    cudaFree(d_all);                                        // on every path
    if (rc != 0) return rc;
    if (!ok) {                                               // NCCL exchange on every rank
        for (u32 r = 0; r < W; ++r) {
            if (r != me && h->peer_ts[r]) cudaIpcCloseMemHandle(h->peer_ts[r]);
            if (r != me && h->peer_flags[r]) cudaIpcCloseMemHandle(h->peer_flags[r]);
            h->peer_ts[r] = h->peer_flags[r] = nullptr;
        }
        cudaGetLastError();
        return 0;
    }
    const size_t off_view = reinterpret_cast<char*>(h->d.view) - reinterpret_cast<char*>(h->d_ts);   // same layout on every rank
    h->p2p_tab = P2PTable{};
    h->p2p_tab.world = W; h->p2p_tab.rank = me;
    for (u32 r = 0; r < W; ++r) {
        h->p2p_tab.slack[r] = reinterpret_cast<u32*>(h->peer_ts[r]);
        h->p2p_tab.view[r] = reinterpret_cast<u64*>(reinterpret_cast<char*>(h->peer_ts[r]) + off_view);
        h->p2p_tab.flags[r] = reinterpret_cast<u64*>(h->peer_flags[r]);
    }
    h->p2p = true;
    return 0;
}

// A bounded spin of the peer-memory exchange expired (a peer never arrived): reported at the next synchronising call.
int p2p_check(abnn_handle* h)
{
    if (!h->p2p) return 0;
    u64 err = 0;
    CU(cudaMemcpyAsync(&err, h->d_p2p + P2P_ERROR, sizeof(u64), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    if (err) return fail(ABNN_ERR_COMM, "peer-memory exchange timed out waiting for a peer rank");
    return 0;
}

// Bring every rank's lastFired slice into the replicated 64-bit snapshot (collective when world_size > 1).
int ensure_view(abnn_handle* h)
{
    if (!h->view_stale) return 0;
    if (!h->comm) return fail(ABNN_ERR_COMM, "world_size > 1 but abnn_comm_init has not been called");
    NC(ncclAllGather(h->d.live + h->lo, h->d.view, h->slice, ncclUint64, h->comm, h->st));
    h->view_stale = false;
    return 0;
}

// After the events of a pass: every rank's owned lastFired slice becomes visible to everyone
// (per-pass allgather of fired-neuron timestamps over NVLink).
// Sharded PARALLEL runs on the line kernel need the remote timestamps only as gate words of the NEXT
// pass, so each rank turns its own slice into slack words (the clock has already been advanced by
// k_end_pass) and the ranks allgather those — half the bytes — plus a broadcast of the input/output
// head of lastFired that teacher forcing and the read-out look at. The 64-bit snapshot of the remote
// slices is refreshed lazily (ensure_view) when something asks for it.
int exchange_timestamps(abnn_handle* h, const KParams& kp)
{
    static const bool full = tune_env("ABNN_FULL_EXCHANGE") != nullptr;       // measurements only
    const u64 head = (u64)h->p.n_input + h->p.n_output;
    h->slack_ready = false;
    h->fire_ready = kp.use_line32 != 0;                      // k_fold_prepare32 has written the owned neurons' words of the next pass
    if (h->p.world_size > 1) {
        if (!h->comm) return fail(ABNN_ERR_COMM, "world_size > 1 but abnn_comm_init has not been called");
        if (slack_mode(h, kp) && h->slice >= head && !full && h->p2p) {
            // peer-memory stores over NVLink instead of the collective (exchange.cu)
            CU(launch_p2p_exchange(kp, h->d, h->p2p_tab, h->lo, h->hi, head, h->sm_count, h->st));
            h->slack_ready = true;
            h->view_stale = true;
        } else if (slack_mode(h, kp) && h->slice >= head && !full) {
            if (!kp.use_line32) CU(launch_build_slack(kp, h->d, h->d.live, h->lo, h->hi, h->st));   // else k_fold_prepare32 built the slice
            NC(ncclGroupStart());
            NC(ncclAllGather(h->d.slack + h->lo, h->d.slack, h->slice, ncclUint32, h->comm, h->st));
            NC(ncclBroadcast(h->d.live, h->d.view, head, ncclUint64, 0, h->comm, h->st));
            NC(ncclGroupEnd());
            h->slack_ready = true;
            h->view_stale = true;
        } else {
            NC(ncclAllGather(h->d.live + h->lo, h->d.view, h->slice, ncclUint64, h->comm, h->st));
            h->view_stale = false;
        }
    } else if (kp.use_line32) {
        h->slack_ready = true;                               // k_fold_prepare32: snapshot carried forward, every gate word rebuilt
    } else if (h->d.view != h->d.live) {
        CU(cudaMemcpyAsync(h->d.view, h->d.live, h->N * sizeof(u64), cudaMemcpyDeviceToDevice, h->st));
    }
    return 0;
}

// Append a chunk of the global table: keep, in order, the records this rank owns.
int append_chunk(abnn_handle* h, const abnn_synapse* host, u64 n)
{
    if (!n) return 0;
    if (h->p.world_size == 1) {
        if (h->n_local + n > h->cap) return fail(ABNN_ERR_CAPACITY, "synapse table capacity exceeded");
        CU(cudaMemcpyAsync(h->d_syn + h->n_local, host, n * sizeof(abnn_synapse), cudaMemcpyHostToDevice, h->st));
        CU(cudaStreamSynchronize(h->st));
        h->n_local += n;
        return 0;
    }
    if (!h->d_stage) {
        h->stage_cap = UPLOAD_CHUNK;
        CU(cudaMalloc(&h->d_stage, h->stage_cap * sizeof(abnn_synapse)));
    }
    for (u64 off = 0; off < n; off += h->stage_cap) {
        const u64 m = std::min(h->stage_cap, n - off);
        CU(cudaMemcpyAsync(h->d_stage, host + off, m * sizeof(abnn_synapse), cudaMemcpyHostToDevice, h->st));
        CompactArgs a{};
        a.in = h->d_stage; a.out = h->d_syn + h->n_local; a.n = m; a.pred = KEEP_OWNED;
        a.dst_lo = (u32)h->lo; a.dst_hi = (u32)h->hi; a.out_cap = h->cap - h->n_local;
        RET(ensure_scratch(h, compact_scratch_bytes(m)));
        CU(launch_compact(a, h->d_scratch, h->d_total, h->st));
        u64 kept = 0;
        CU(cudaMemcpyAsync(&kept, h->d_total, sizeof(u64), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        if (kept > h->cap - h->n_local) return fail(ABNN_ERR_CAPACITY, "synapse table capacity exceeded (owned share larger than syn_capacity)");
        h->n_local += kept;
    }
    return 0;
}

// A table that came from the host (upload, .bnn): the kernels index the per-neuron arrays with src and dst of every record,
// so a record that names a neuron the handle does not have is refused here, and the handle is left with an empty table.
int reset_structural(abnn_handle* h);
int check_table(abnn_handle* h, bool allow_dead)
{
    u64 bad = 0;
    CU(launch_validate_table(h->d_syn, h->n_local, (u32)h->N, (u32)h->lo, (u32)h->hi, allow_dead, h->d_total + 3, h->sm_count, h->st));
    CU(cudaMemcpyAsync(&bad, h->d_total + 3, sizeof(u64), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    if (!bad) return 0;
    h->n_local = 0; h->counts_dirty = true;
    if (h->p.world_size == 1) h->n_local_all.assign(1, 0);
    reset_structural(h);
    return fail(ABNN_ERR_INVALID, std::to_string(bad) + " synapse record(s) name a neuron outside the handle's " + std::to_string(h->N) +
                                  " neurons (src, or dst outside this rank's slice); the table was dropped");
}

// A new table (upload / init / load): the structural-step bookkeeping starts over — the whole table is the ordered region.
int reset_structural(abnn_handle* h)
{
    h->struct_steps = 0; h->n_sorted = h->n_local; h->n_dead = 0;
    k_set_struct<<<1, 1, 0, h->st>>>(h->d.sc, 0, h->n_local, 0);
    CU(cudaGetLastError());
    return 0;
}

// ABNN_TABLE_DST_SORTED: keep the rank's table stably sorted by destination neuron (include/abnn.h).
// Scratch (a second table + two key arrays) is allocated for the duration of the sort only.
int sort_table(abnn_handle* h)
{
    if (h->p.table_order == ABNN_TABLE_AS_GIVEN || h->n_local < 2) return 0;
    const u64 n = h->n_local;
    const bool interleave = h->p.table_order == ABNN_TABLE_DST_INTERLEAVED;
    abnn_synapse* alt = nullptr; u32* keys = nullptr; void* tmp = nullptr; u64* starts = nullptr;
    const size_t tmp_bytes = sort_by_dst_temp_bytes(n);
    auto release = [&] { cudaFree(alt); cudaFree(keys); cudaFree(tmp); cudaFree(starts); };
    cudaError_t e = cudaMalloc(&alt, n * sizeof(abnn_synapse));
    if (e == cudaSuccess) e = cudaMalloc(&keys, 2 * n * sizeof(u32));
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
    if (e == cudaSuccess && interleave) e = cudaMalloc(&starts, (h->hi - h->lo + 1) * sizeof(u64));
    bool in_alt = false;
    int nb = 1; while ((1ull << nb) < h->N) ++nb;
    if (e == cudaSuccess) e = launch_sort_by_dst(h->d_syn, alt, keys, keys + n, n, nb, tmp, tmp_bytes, &in_alt, h->st);
    if (e == cudaSuccess && interleave) {                    // sorted table -> interleaved, into the other buffer
        e = launch_interleave_by_dst(in_alt ? alt : h->d_syn, in_alt ? h->d_syn : alt, n, (u32)h->lo, (u32)(h->hi - h->lo),
                                     h->p.sample_block >= 16 ? 16u : 8u, starts, h->st);
        in_alt = !in_alt;
    }
    if (e == cudaSuccess && in_alt) e = cudaMemcpyAsync(h->d_syn, alt, n * sizeof(abnn_synapse), cudaMemcpyDeviceToDevice, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    release();
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ABNN_ERR_CUDA, std::string("sort_table: ") + cudaGetErrorString(e)); }
    return 0;
}

// ABNN_TABLE_DST_SORTED growth: the m candidates list[0..m) (in append order) become records, are sorted by dst
// (stable) and merged into the sorted table out of place; the tables are then swapped (scratch and the spare table
// are cached in the handle). With kept_out != null the
// pruning (w < w_prune) happens in the same pass over the table (launch_prune_merge_sorted) and *kept_out receives the
// number of existing records that survived.
// With list == null the m new records are already in h->mg_new (the filtered tail of a periodic rebuild) and n_ordered is
// the size of the ordered region they are merged into.
// Scratch for m new records (two copies + sort keys), cached in the handle: cudaMalloc / cudaFree next to a 16 GB table
// cost more than the merge itself (18 - 78 ms measured), so it is sized by `expect` (what later steps will need) when that
// is larger and never shrinks.
static int ensure_merge_scratch(abnn_handle* h, u64 m, u64 expect)
{
    if (m <= h->mg_cap) return 0;
    const u64 want = std::min<u64>(std::max(m, expect), 1ull << 31);
    cudaFree(h->mg_new); cudaFree(h->mg_keys); cudaFree(h->mg_tmp);
    h->mg_new = nullptr; h->mg_keys = nullptr; h->mg_tmp = nullptr; h->mg_cap = 0;
    const u32 cap_m = next_pow2((u32)want);
    h->mg_tmp_bytes = sort_by_dst_temp_bytes(cap_m);
    CU(cudaMalloc(&h->mg_new, 2 * (size_t)cap_m * sizeof(abnn_synapse)));
    CU(cudaMalloc(&h->mg_keys, 2 * (size_t)cap_m * sizeof(u32)));
    CU(cudaMalloc(&h->mg_tmp, h->mg_tmp_bytes ? h->mg_tmp_bytes : 16));
    h->mg_cap = cap_m;
    return 0;
}

int merge_grown(abnn_handle* h, const GrowCand* list, u32 m, u64* kept_out = nullptr, u64 n_ordered = ~0ull)
{
    const u64 n = n_ordered == ~0ull ? h->n_local : n_ordered;
    const u32 span = (u32)(h->hi - h->lo);
    if (m > h->mg_cap) {
        if (!list) return fail(ABNN_ERR_INVALID, "merge_grown: ready records need the scratch sized beforehand");
        RET(ensure_merge_scratch(h, m, 2ull * m));
    }
    if (!h->mg_cnt) {
        h->mg_scan_bytes = merge_scan_temp_bytes((u64)span + 2);
        CU(cudaMalloc(&h->mg_cnt, ((size_t)span + 1) * sizeof(u32)));
        CU(cudaMalloc(&h->mg_scan, h->mg_scan_bytes ? h->mg_scan_bytes : 16));
    }
    if (kept_out) {
        if (!h->mg_pruned) CU(cudaMalloc(&h->mg_pruned, ((size_t)span + 2) * sizeof(u32)));
        RET(ensure_scratch(h, std::max(compact_scratch_bytes(h->cap), compact2_scratch_bytes(h->cap))));   // by capacity: never regrown
    }
    abnn_synapse* out = h->d_spare;
    h->d_spare = nullptr;
    if (!out) CU(cudaMalloc(&out, h->cap * sizeof(abnn_synapse)));
    abnn_synapse *nw = h->mg_new, *nw_alt = h->mg_new + h->mg_cap;
    cudaError_t e = cudaMemsetAsync(h->mg_cnt, 0, ((size_t)span + 1) * sizeof(u32), h->st);
    if (e == cudaSuccess && list) e = launch_grow_append(list, m, nw, 0, h->p.w_init, h->st);
    bool in_alt = false;
    int nb = 1; while ((1ull << nb) < h->N) ++nb;
    if (e == cudaSuccess) e = launch_sort_by_dst(nw, nw_alt, h->mg_keys, h->mg_keys + h->mg_cap, m, nb, h->mg_tmp, h->mg_tmp_bytes, &in_alt, h->st);
    u64 kept = n;
    if (kept_out) {
        if (e == cudaSuccess) e = cudaMemsetAsync(h->mg_pruned, 0, ((size_t)span + 2) * sizeof(u32), h->st);
        if (e == cudaSuccess) e = launch_prune_merge_sorted(h->d_syn, n, h->p.w_prune, in_alt ? nw_alt : nw, m, (u32)h->lo, span, h->mg_cnt,
                                                            h->mg_pruned, h->mg_scan, h->mg_scan_bytes, h->d_scratch, h->d_total, out,
                                                            h->cap, h->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&kept, h->d_total, sizeof(u64), cudaMemcpyDeviceToHost, h->st);
    } else if (e == cudaSuccess) {
        e = launch_merge_sorted(h->d_syn, n, in_alt ? nw_alt : nw, m, (u32)h->lo, span, h->mg_cnt, h->mg_scan, h->mg_scan_bytes, out,
                                h->sm_count, h->st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e != cudaSuccess) {
        cudaFree(out); cudaGetLastError();
        return fail(ABNN_ERR_CUDA, std::string("merge_grown: ") + cudaGetErrorString(e));
    }
    // the old table becomes the spare of the next growth step (freeing and re-allocating 16 GB costs ~18 ms)
    // unless two tables would take more than half of the device memory
    if (2 * h->cap * sizeof(abnn_synapse) <= h->mem_total / 2) h->d_spare = h->d_syn;
    else cudaFree(h->d_syn);
    h->d_syn = out; h->d.syn = out;
    if (h->step_exec) { cudaGraphExecDestroy(h->step_exec); h->step_exec = nullptr; }   // the captured pass holds the old table pointer
    h->n_local = kept + m;
    if (kept_out) *kept_out = kept;
    return 0;
}

// EXACT execution: phase 1 (open events, counted per destination) -> scan + scatter into per-destination buckets ->
// phase 3 (per-destination chains). No count leaves the device: the buffers are sized by the events of the pass and the
// kernels read the counts from device memory — no host synchronisation inside a pass, so an EXACT pass can be enqueued
// asynchronously and captured into a CUDA graph like a PARALLEL one.
int ensure_exact_scratch(abnn_handle* h, const KParams& kp)
{
    if (kp.count >= (1ull << 32)) return fail(ABNN_ERR_UNSUPPORTED, "EXACT execution: at most 2^32-1 events per rank per pass");
    const u64 span = h->hi - h->lo;
    if (!h->d_xcnt) {
        CU(cudaMalloc(&h->d_xcnt, 2 * std::max<u64>(span, 1) * sizeof(u32)));            // counts, then cursors
        CU(cudaMalloc(&h->d_xcount, 2 * sizeof(u32)));
        h->x_tmp_bytes = exact_scan_temp_bytes(span);
        CU(cudaMalloc(&h->d_xtmp, h->x_tmp_bytes ? h->x_tmp_bytes : 16));
    }
    if (kp.count > h->x_cap) {
        cudaFree(h->d_xkeys); cudaFree(h->d_xbucket);
        h->d_xkeys = nullptr; h->d_xbucket = nullptr; h->x_cap = 0;
        const u64 cap = kp.count;
        CU(cudaMalloc(&h->d_xkeys, cap * sizeof(u64)));
        CU(cudaMalloc(&h->d_xbucket, 2 * cap * sizeof(u32)));                            // buckets, then arrival numbers
        h->x_cap = cap;
        if (h->step_exec) { cudaGraphExecDestroy(h->step_exec); h->step_exec = nullptr; }   // the captured pass holds the old buffers
    }
    return 0;
}
int run_exact(abnn_handle* h, const KParams& kp)
{
    RET(ensure_exact_scratch(h, kp));                 // no-op once the buffers fit (abnn_engine_step sizes them before a capture)
    if (!kp.count) return 0;
    const u32 lo = (u32)h->lo, span = (u32)(h->hi - h->lo);
    u32 *cnt = h->d_xcnt, *cursor = h->d_xcnt + std::max<u64>(span, 1);
    u32* slot = h->d_xbucket + h->x_cap;
    CU(launch_exact_phase1(kp, h->d, h->d_xkeys, slot, cnt, lo, span, h->d_xcount, h->sm_count, h->st));
    CU(launch_exact_group(h->d_xkeys, slot, h->d_xcount, cnt, cursor, lo, span, h->d_xbucket, h->d_xtmp, h->x_tmp_bytes, h->sm_count, h->st));
    CU(launch_exact_phase3(kp, h->d, h->d_xbucket, cnt, cursor, lo, span, h->d_xcount, h->sm_count, h->st));
    return 0;
}

}  // namespace

// The persisting-L2 set-aside is a DEVICE-wide limit: the handles that need it share it. The first one remembers what the
// limit was, the last one to go puts it back (a set-aside nobody fills is L2 taken away from everything else that runs on
// the device, e.g. a later EXACT handle: 9.5 against 11.0 ms per pass).
namespace {
std::mutex g_l2_mu;
int g_l2_users[64] = {};
size_t g_l2_before[64] = {};
}  // namespace
void l2_acquire(abnn_handle* h, size_t before)
{
    if (h->device < 0 || h->device >= 64 || h->l2_user) return;
    std::lock_guard<std::mutex> lk(g_l2_mu);
    if (g_l2_users[h->device]++ == 0) g_l2_before[h->device] = before;
    h->l2_user = true;
}
void l2_release(abnn_handle* h)
{
    if (!h->l2_user) return;
    std::lock_guard<std::mutex> lk(g_l2_mu);
    h->l2_user = false;
    if (--g_l2_users[h->device] == 0) {
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, g_l2_before[h->device]);
        cudaGetLastError();
    }
}

// =================================================================================================
extern "C" {

const char* abnn_last_error(void) { return g_err.c_str(); }
uint32_t abnn_abi_version(void) { return ABNN_ABI_VERSION; }

int abnn_default_params(abnn_params* p, uint32_t profile)
{
    if (!p) return fail(ABNN_ERR_INVALID, "null params");
    if (profile > ABNN_PROFILE_B200) return fail(ABNN_ERR_INVALID, "unknown profile");
    std::memset(p, 0, sizeof(*p));
    p->struct_size = sizeof(abnn_params); p->abi_version = ABNN_ABI_VERSION;
    p->n_input = 256; p->n_output = 256; p->n_hidden = 5000000ull; p->n_syn = 1000000000ull;   // constants.h:2-5
    p->seed = 42;                                                                               // simple.yml:12
    if (profile == ABNN_PROFILE_METAL_PARITY) {
        p->sampler = ABNN_SAMPLER_SWEEP; p->release_rng = ABNN_RNG_XORSHIFT;
        p->clock_mode = ABNN_CLOCK_PER_PASS; p->exec_mode = ABNN_EXEC_SERIAL;
        p->src_view = ABNN_SRC_LIVE; p->rbar_mode = ABNN_RBAR_METAL_TID0;
        p->max_spikes_per_pass = 2560; p->track_visits = 0;                                     // brain.h:18
        p->window_pre = 5; p->refractory = 2;                                                   // brain.metal:23-24
    } else {
        p->sampler = ABNN_SAMPLER_PHILOX; p->release_rng = ABNN_RNG_PHILOX;
        p->clock_mode = ABNN_CLOCK_PER_EVENT; p->exec_mode = ABNN_EXEC_PARALLEL;
        p->src_view = ABNN_SRC_SNAPSHOT; p->rbar_mode = ABNN_RBAR_PASS_STEP;
        p->max_spikes_per_pass = 0; p->track_visits = 1;
        p->window_pre = 50000; p->refractory = 2;                                               // brain.cpp:102
    }
    p->teacher_gap = 1;                                                                         // brain-engine.cpp:130
    p->base_scale = 0.8f; p->a_ltp = 0.04f; p->a_ltd = 0.02f; p->w_min = 0.001f; p->w_max = 1.0f;
    p->eta_home = 1.0e-6f; p->target_rate_hz = 1000.0f; p->home_tick_hz = 1e6f;
    p->eta_reward = 1.0e-3f; p->alpha_rbar = 0.001f;
    p->w_prune = 0.f; p->p_new = 0.f; p->w_init = 0.1f;
    p->rate_alpha = 0.5f; p->peak_decay = 0.999f; p->peak_init = 0.5f;
    p->use_fir = 1; p->fir_size = 20; p->reward_window = 1000;
    p->filter_tau = 0.02; p->dt_sec = 0.0009; p->loss0 = 0.25;
    p->device = -1; p->rank = 0; p->world_size = 1; p->l2_persist = 1;
    p->sample_block = 1;
    if (profile == ABNN_PROFILE_B200) { p->sample_block = 16; p->table_order = ABNN_TABLE_DST_INTERLEAVED; }
    return 0;
}

int abnn_partition(uint64_t n_neuron, uint32_t world, uint32_t rank, uint64_t* lo, uint64_t* hi)
{
    if (!world || rank >= world || !lo || !hi) return fail(ABNN_ERR_INVALID, "bad partition arguments");
    const u64 slice = (n_neuron + world - 1) / world;
    *lo = std::min<u64>(n_neuron, slice * rank);
    *hi = std::min<u64>(n_neuron, slice * (rank + 1));
    return 0;
}
int abnn_event_share(uint64_t events, uint64_t n_global, uint64_t before, uint64_t n_local, uint64_t* first, uint64_t* count)
{
    if (!first || !count) return fail(ABNN_ERR_INVALID, "null output");
    u64 f, c; event_share(events, n_global, before, n_local, &f, &c);
    *first = f; *count = c;
    return 0;
}
void abnn_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// ---- lifetime -------------------------------------------------------------------------------------
int abnn_create(const abnn_params* pp, abnn_handle** out)
{
    if (!pp || !out) return fail(ABNN_ERR_INVALID, "null argument");
    *out = nullptr;
    const abnn_params& p = *pp;
    if (p.struct_size != sizeof(abnn_params) || p.abi_version != ABNN_ABI_VERSION)
        return fail(ABNN_ERR_INVALID, "abnn_params size/version mismatch (header and library differ)");
    if (!p.world_size || p.rank >= p.world_size) return fail(ABNN_ERR_INVALID, "bad rank/world_size");
    if (p.sampler > 1 || p.release_rng > 1 || p.clock_mode > 1 || p.exec_mode > 2 || p.src_view > 1 || p.rbar_mode > 1)
        return fail(ABNN_ERR_INVALID, "unknown mode value");
    if (p.table_order > ABNN_TABLE_DST_INTERLEAVED) return fail(ABNN_ERR_INVALID, "unknown table_order");
    if (p.exchange > ABNN_EXCHANGE_PEER || p.prune_in_place > 1) return fail(ABNN_ERR_INVALID, "unknown exchange / prune_in_place value");
    if (p.compact_every > 1 && p.w_prune > 0.f && p.p_new > 0.f && p.w_init < p.w_prune)
        return fail(ABNN_ERR_INVALID, "compact_every > 1 needs w_init >= w_prune (a grown synapse below the pruning threshold would never be staged)");
    if (p.sample_block > 32 || (p.sample_block & (p.sample_block - 1)))
        return fail(ABNN_ERR_INVALID, "sample_block must be a power of two <= 32 (0 = 1)");
    if (!p.n_output || p.fir_size == 0 || p.fir_size > ABNN_MAX_FIR) return fail(ABNN_ERR_INVALID, "bad n_output / fir_size");
    const u64 N = (u64)p.n_input + p.n_output + p.n_hidden;
    if (N >= (1ull << 32)) return fail(ABNN_ERR_INVALID, "neuron ids are 32-bit (SynapsePacked.src/dst)");
    if (p.src_view == ABNN_SRC_LIVE && p.world_size > 1)
        return fail(ABNN_ERR_UNSUPPORTED, "LIVE src view is single-GPU only; sharded runs read the pass-start snapshot");
    if (p.rbar_mode == ABNN_RBAR_METAL_TID0 && p.exec_mode != ABNN_EXEC_SERIAL)
        return fail(ABNN_ERR_UNSUPPORTED, "METAL_TID0 r-bar needs SERIAL execution");
    if (p.exec_mode == ABNN_EXEC_EXACT && p.src_view != ABNN_SRC_SNAPSHOT)
        return fail(ABNN_ERR_UNSUPPORTED, "EXACT execution needs the SNAPSHOT src view");
    if (p.exec_mode == ABNN_EXEC_EXACT && p.max_spikes_per_pass != 0)
        return fail(ABNN_ERR_UNSUPPORTED, "EXACT execution does not support the ordered spike budget (max_spikes_per_pass must be 0)");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ABNN_ERR_NO_DEVICE, "no CUDA device: abnn_b200 has no CPU fallback");
    }
    int dev = p.device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(ABNN_ERR_INVALID, "device ordinal out of range");
    CU(cudaSetDevice(dev));

    abnn_handle* h = new abnn_handle;
    h->p = p; h->device = dev; h->N = N;
    h->slice = (N + p.world_size - 1) / p.world_size;
    h->lo = std::min<u64>(N, h->slice * p.rank);
    h->hi = std::min<u64>(N, h->slice * (p.rank + 1));
    h->npad = h->slice * p.world_size;
    h->n_local_all.assign(p.world_size, 0);
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); h->sm_count = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev); h->l2_bytes = (size_t)v;
    { size_t mem_free = 0; cudaMemGetInfo(&mem_free, &h->mem_total); }
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);

    h->cap = p.syn_capacity;
    if (!h->cap) {
        h->cap = p.world_size == 1 ? p.n_syn
                                   : (p.n_syn + p.world_size - 1) / p.world_size + p.n_syn / (16ull * p.world_size) +
                                         65536 + (u64)p.n_input * p.n_output;
    }
    if (!h->cap) h->cap = 1;

#define CUH(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
            abnn_destroy(h);                                                                           \
            return fail(ABNN_ERR_CUDA, m_);                                                            \
        }                                                                                              \
    } while (0)

    CUH(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    CUH(cudaEventCreate(&h->ev0));
    CUH(cudaEventCreate(&h->ev1));
    CUH(cudaEventCreate(&h->evj));
    CUH(cudaEventCreate(&h->evk));
    for (int i = 0; i < 8; ++i) CUH(cudaEventCreate(&h->timer[i]));
    for (int i = 0; i < RING; ++i) CUH(cudaEventCreateWithFlags(&h->ring_ev[i], cudaEventDisableTiming));
    CUH(cudaMalloc(&h->d_syn, h->cap * sizeof(abnn_synapse)));
    // timestamps in one allocation so that one access-policy window covers the hot arrays, hottest first:
    //   SNAPSHOT view: [slack32 | fire32 | vis32 | visited | live | view]   LIVE view: [live | visited]
    // slack32 / fire32 / vis32 are the per-pass 32-bit forms of the snapshot, lastFired and lastVisited that the line
    // kernel works on (traversal.cu:k_traverse_line32, k_prepare32, k_fold32): 12 bytes per neuron = 60 MB at 5M neurons,
    // the size of the persisting-L2 set-aside. The 64-bit arrays serve every other kernel, the read-out and the ABI.
    // fire32 / vis32 exist for the OWNED neurons only (a sharded rank touches nothing else) and are indexed by global id
    // through pointers offset by -lo: the persisting set-aside is then exactly what the rank touches — 4 N + 8 N/G bytes.
    // (Measured at G = 2 with full-size arrays: the 60 MB set-aside, a third of it never touched, made the pass SLOWER
    // than no set-aside at all, 0.84 vs 0.76 ms: profiles/r2_notes.md §5.)
    const u64 npad = (h->npad + 31) & ~31ull;
    const u64 spad = (h->slice + 31) & ~31ull;
    const bool snap = p.src_view == ABNN_SRC_SNAPSHOT;
    const u64 words = snap ? 7 * npad + 2 * spad : 4 * npad;   // in units of 4 bytes
    CUH(cudaMalloc(&h->d_ts, words * sizeof(u32)));
    CUH(cudaMemsetAsync(h->d_ts, 0, words * sizeof(u32), h->st));                               // brain.cpp:62-64
    if (snap) {
        h->d.slack = reinterpret_cast<u32*>(h->d_ts);
        h->d.fire32 = reinterpret_cast<int*>(h->d.slack + npad) - h->lo;
        h->d.vis32 = h->d.slack + npad + spad - h->lo;
        h->d.visited = reinterpret_cast<u64*>(h->d.slack + npad + 2 * spad);
        h->d.live = h->d.visited + npad;
        h->d.view = h->d.live + npad;
    } else {
        h->d.slack = nullptr; h->d.fire32 = nullptr; h->d.vis32 = nullptr;
        h->d.live = h->d_ts; h->d.view = h->d_ts;
        h->d.visited = h->d_ts + npad;
    }
    h->d.syn = h->d_syn;
    CUH(cudaMalloc(&h->d.sc, sizeof(DevScalars)));
    {
        DevScalars s{};
        s.max_observed = p.peak_init; s.last_loss = p.loss0; s.last_pass_ticks = 1;              // brain-engine.h:54,83
        CUH(cudaMemcpy(h->d.sc, &s, sizeof(s), cudaMemcpyHostToDevice));
    }
    h->grow_cap = GROW_CAP_DEFAULT;
    h->grow_buf = next_pow2(h->grow_cap);
    if (p.p_new > 0.f) CUH(cudaMalloc(&h->d.grow, (size_t)h->grow_buf * sizeof(GrowCand)));
    if (p.compact_every > 1 && p.w_prune > 0.f) CUH(cudaMalloc(&h->d.prune_list, (size_t)PRUNE_CAP_DEFAULT * sizeof(u64)));
    h->vec_len = std::max(p.n_input, p.n_output);
    CUH(cudaMalloc(&h->d_vec, (size_t)RING * h->vec_len * sizeof(float)));
    CUH(cudaMallocHost(&h->h_vec, (size_t)RING * h->vec_len * sizeof(float)));
    CUH(cudaMalloc(&h->rs.rate, p.n_output * sizeof(float)));
    CUH(cudaMalloc(&h->rs.iir, p.n_output * sizeof(float)));
    CUH(cudaMalloc(&h->rs.fir, (size_t)p.fir_size * p.n_output * sizeof(float)));
    CUH(cudaMalloc(&h->rs.smooth, p.n_output * sizeof(float)));
    CUH(cudaMalloc(&h->rs.spikes, p.n_output));
    CUH(cudaMemsetAsync(h->rs.rate, 0, p.n_output * sizeof(float), h->st));
    CUH(cudaMemsetAsync(h->rs.iir, 0, p.n_output * sizeof(float), h->st));
    CUH(cudaMemsetAsync(h->rs.fir, 0, (size_t)p.fir_size * p.n_output * sizeof(float), h->st));
    CUH(cudaMemsetAsync(h->rs.smooth, 0, p.n_output * sizeof(float), h->st));
    CUH(cudaMalloc(&h->d_stats, sizeof(abnn_pass_stats)));
    CUH(cudaMemsetAsync(h->d_stats, 0, sizeof(abnn_pass_stats), h->st));
    h->h_pin_bytes = std::max<size_t>(4096, (size_t)p.n_output * sizeof(float) * 2 + sizeof(DevScalars));
    CUH(cudaMallocHost(&h->h_pin, h->h_pin_bytes));
    CUH(cudaMalloc(&h->d_total, 8 * sizeof(u64)));
    CUH(cudaMalloc(&h->d_counts, p.world_size * sizeof(u64)));

    // L2 residency of the timestamp arrays (north star item 2): params.l2_persist = 1 sets aside persisting L2 for the hot
    // arrays (cudaLimitPersistingL2CacheSize is a DEVICE-wide limit: the handle raises it to what it needs and never lowers
    // it) and attaches an access-policy window to the handle's stream; l2_persist = 0 touches no device-wide state.
    // Tuning builds: ABNN_L2_ARRAYS = arrays covered by the window, ABNN_L2_MISS = 1 -> lines of the window that do not get
    // the persisting property are "normal" instead of "streaming".
    // Only PARALLEL execution reads the arrays the window covers; a set-aside nobody fills is L2 taken away from the 64-bit
    // arrays EXACT / SERIAL execution work on (measured: 11.0 -> 9.5 ms per EXACT pass at the 1B shape without it).
    if (p.l2_persist && p.exec_mode == ABNN_EXEC_PARALLEL && max_persist > 0 && max_window > 0) {
        const int n_arr = tune_env("ABNN_L2_ARRAYS") ? atoi(tune_env("ABNN_L2_ARRAYS")) : 3;
        const bool miss_normal = tune_env("ABNN_L2_MISS") && atoi(tune_env("ABNN_L2_MISS")) == 1;
        // default: slack32 + the owned fire32 / vis32 (60 MB at 5M neurons on one GPU) / LIVE view: lastFired + lastVisited
        size_t hot = snap ? (size_t)(npad + (n_arr >= 2 ? spad : 0) + (n_arr >= 3 ? spad : 0)) * sizeof(u32)
                          : (size_t)(n_arr <= 1 ? 2 : 4) * npad * sizeof(u32);
        size_t want = std::min<size_t>(hot, (size_t)max_persist);
        if (tune_env("ABNN_L2_CARVE_MAX")) want = (size_t)max_persist;
        size_t before = 0;
        if (cudaDeviceGetLimit(&before, cudaLimitPersistingL2CacheSize) != cudaSuccess) { before = 0; cudaGetLastError(); }
        if (before > want) want = before;
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            l2_acquire(h, before);
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = h->d_ts;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(hot, (size_t)max_window);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)got / (double)attr.accessPolicyWindow.num_bytes);
            if (tune_env("ABNN_L2_RATIO")) attr.accessPolicyWindow.hitRatio = (float)atof(tune_env("ABNN_L2_RATIO"));
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = miss_normal ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(h->st, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess) h->l2_persist = got;
        }
        cudaGetLastError();
    }
    CUH(cudaStreamSynchronize(h->st));
#undef CUH
    *out = h;
    return 0;
}

void abnn_destroy(abnn_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    l2_release(h);
    for (u32 r = 0; r < P2P_MAX_WORLD; ++r) {
        if (h->p2p && r != h->p.rank && h->peer_ts[r]) cudaIpcCloseMemHandle(h->peer_ts[r]);
        if (h->p2p && r != h->p.rank && h->peer_flags[r]) cudaIpcCloseMemHandle(h->peer_flags[r]);
    }
    cudaFree(h->d_p2p);
    if (h->comm) ncclCommDestroy(h->comm);
    cudaFree(h->d_syn); cudaFree(h->d_spare); cudaFree(h->mg_new); cudaFree(h->mg_keys); cudaFree(h->mg_tmp); cudaFree(h->mg_cnt);
    cudaFree(h->mg_pruned); cudaFree(h->mg_scan); cudaFree(h->d_ts); cudaFree(h->d.sc); cudaFree(h->d.grow); cudaFree(h->d_grow_all);
    cudaFree(h->d.prune_list);
    cudaFree(h->d_vec); cudaFreeHost(h->h_vec);
    cudaFree(h->rs.rate); cudaFree(h->rs.iir); cudaFree(h->rs.fir); cudaFree(h->rs.smooth); cudaFree(h->rs.spikes);
    cudaFree(h->d_stats); cudaFreeHost(h->h_pin); cudaFree(h->d_scratch); cudaFree(h->d_stage);
    cudaFree(h->d_total); cudaFree(h->d_counts);
    if (h->step_exec) cudaGraphExecDestroy(h->step_exec);
    cudaFree(h->d_frame); cudaFreeHost(h->h_frame);
    for (int i = 0; i < RING; ++i) if (h->frame_ev[i]) cudaEventDestroy(h->frame_ev[i]);
    cudaFree(h->d_xkeys); cudaFree(h->d_xbucket); cudaFree(h->d_xcnt); cudaFree(h->d_xcount); cudaFree(h->d_xtmp);
    for (int i = 0; i < RING; ++i) if (h->ring_ev[i]) cudaEventDestroy(h->ring_ev[i]);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evj) cudaEventDestroy(h->evj);
    if (h->evk) cudaEventDestroy(h->evk);
    for (int i = 0; i < 8; ++i) if (h->timer[i]) cudaEventDestroy(h->timer[i]);
    if (h->st) cudaStreamDestroy(h->st);
    cudaGetLastError();
    delete h;
}

int abnn_get_info(abnn_handle* h, abnn_info* o)
{
    RET(use(h));
    if (!o) return fail(ABNN_ERR_INVALID, "null output");
    RET(refresh_counts(h));
    DevScalars s; RET(read_scalars(h, &s));
    std::memset(o, 0, sizeof(*o));
    o->n_input = h->p.n_input; o->n_output = h->p.n_output; o->n_hidden = h->p.n_hidden; o->n_neuron = h->N;
    for (u64 v : h->n_local_all) o->n_syn_global += v;
    o->n_syn_local = h->n_local; o->syn_capacity = h->cap;
    o->neuron_lo = h->lo; o->neuron_hi = h->hi; o->neuron_slice = h->slice;
    o->rank = h->p.rank; o->world_size = h->p.world_size; o->device = h->device; o->sm_count = h->sm_count;
    o->l2_bytes = h->l2_bytes; o->l2_persist_bytes = h->l2_persist;
    o->pass_index = s.pass_index; o->clock = s.clock; o->event_base = s.event_base;
    return 0;
}

// ---- communicator ---------------------------------------------------------------------------------
int abnn_comm_unique_id(void* id128)
{
    if (!id128) return fail(ABNN_ERR_INVALID, "null id buffer");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(ncclGetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return 0;
}
int abnn_comm_init(abnn_handle* h, const void* id128)
{
    RET(use(h));
    if (!id128) return fail(ABNN_ERR_INVALID, "null id");
    if (h->comm) return fail(ABNN_ERR_INVALID, "communicator already initialised");
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    NC(ncclCommInitRank(&h->comm, (int)h->p.world_size, id, (int)h->p.rank));
    if (h->p.exchange == ABNN_EXCHANGE_PEER) RET(p2p_setup(h));   // peer-memory exchange instead of the NCCL allgather
    return 0;
}

// ---- graph ----------------------------------------------------------------------------------------
int abnn_upload_synapses(abnn_handle* h, const abnn_synapse* syn, uint64_t n)
{
    RET(use(h));
    if (!syn && n) return fail(ABNN_ERR_INVALID, "null table");
    h->n_local = 0;
    RET(append_chunk(h, syn, n));
    RET(check_table(h, false));
    RET(sort_table(h));
    RET(reset_structural(h));
    h->counts_dirty = true;
    if (h->p.world_size == 1) h->n_local_all.assign(1, h->n_local);
    return 0;
}

int abnn_download_synapses(abnn_handle* h, abnn_synapse* out, uint64_t cap, uint64_t* n_out)
{
    RET(use(h));
    if (n_out) *n_out = h->n_local;
    if (cap < h->n_local) return fail(ABNN_ERR_CAPACITY, "output buffer smaller than the live table");
    if (h->n_local && !out) return fail(ABNN_ERR_INVALID, "null output");
    CU(cudaMemcpyAsync(out, h->d_syn, h->n_local * sizeof(abnn_synapse), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

int abnn_init_graph(abnn_handle* h, uint32_t kind, uint64_t seed)
{
    RET(use(h));
    const abnn_params& p = h->p;
    if (kind == ABNN_GRAPH_ER_BETA) {
        const u64 g0 = (u64)((unsigned __int128)p.n_syn * p.rank / p.world_size);
        const u64 g1 = (u64)((unsigned __int128)p.n_syn * (p.rank + 1) / p.world_size);
        if (g1 - g0 > h->cap) return fail(ABNN_ERR_CAPACITY, "synapse table capacity exceeded");
        if (h->hi == h->lo && g1 > g0) return fail(ABNN_ERR_INVALID, "rank owns no neurons");
        CU(launch_init_er_beta(h->d_syn, g0, g1 - g0, seed, h->N, h->lo, h->hi, h->sm_count, h->st));
        h->n_local = g1 - g0;
        RET(sort_table(h));
        RET(reset_structural(h));
        for (u32 k = 0; k < p.world_size; ++k)
            h->n_local_all[k] = (u64)((unsigned __int128)p.n_syn * (k + 1) / p.world_size) -
                                (u64)((unsigned __int128)p.n_syn * k / p.world_size);
        h->counts_dirty = false;
        return 0;
    }
    if (kind == ABNN_GRAPH_REFERENCE) {
        // build_random_graph (brain-engine.cpp:31-53) streamed in chunks: mt19937(seed) [the reference
        // seeds with 1], dense input->output w~U[.4,.8), then hidden->hidden w~U[.1,.2) with the
        // draw order hid(src), hid(dst), wHH. Uses the host C++ library's distributions, as the
        // reference does, so the table equals what the reference builds with this toolchain.
        if (p.n_hidden == 0 && p.n_syn > (u64)p.n_input * p.n_output) return fail(ABNN_ERR_INVALID, "no hidden neurons for hidden->hidden edges");
        std::mt19937 gen((uint32_t)seed);
        std::uniform_real_distribution<float> wIn(0.4f, 0.8f), wHH(0.1f, 0.2f);
        std::uniform_int_distribution<uint32_t> hid(p.n_input + p.n_output, (uint32_t)(h->N - 1));
        const u64 chunk = std::min<u64>(UPLOAD_CHUNK, std::max<u64>(p.n_syn, 1));
        std::vector<abnn_synapse> buf(chunk);
        h->n_local = 0;
        u64 idx = 0, fill = 0;
        uint32_t i = 0, o = 0;
        const u64 dense = std::min<u64>(p.n_syn, (u64)p.n_input * p.n_output);
        while (idx < p.n_syn) {
            if (idx < dense) {
                buf[fill++] = abnn_synapse{i, p.n_input + o, wIn(gen), 0.f};
                if (++o == p.n_output) { o = 0; ++i; }
            } else {
                const uint32_t a = hid(gen);
                const uint32_t b = hid(gen);
                const float    w = wHH(gen);
                buf[fill++] = abnn_synapse{a, b, w, 0.f};
            }
            ++idx;
            if (fill == chunk || idx == p.n_syn) { RET(append_chunk(h, buf.data(), fill)); fill = 0; }
        }
        RET(sort_table(h));
        RET(reset_structural(h));
        h->counts_dirty = true;
        if (p.world_size == 1) h->n_local_all.assign(1, h->n_local);
        return 0;
    }
    return fail(ABNN_ERR_INVALID, "unknown graph kind");
}

// .bnn v1 (brain.cpp:161-178): u32 N_SYN, u32 N_NRN, N_SYN x 16 B, no padding.
int abnn_save_bnn(abnn_handle* h, const char* path)
{
    RET(use(h));
    if (!path) return fail(ABNN_ERR_INVALID, "null path");
    if (h->p.world_size != 1) return fail(ABNN_ERR_UNSUPPORTED, ".bnn v1 holds one unsharded table; save from a single-GPU handle");
    if (h->n_local >= (1ull << 32)) return fail(ABNN_ERR_UNSUPPORTED, ".bnn v1 header counts are 32-bit");
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(ABNN_ERR_IO, std::string("cannot open for writing: ") + path);
    const uint32_t hdr[2] = {(uint32_t)h->n_local, (uint32_t)h->N};
    bool ok = std::fwrite(hdr, 4, 2, f) == 2;
    const u64 chunk = 4ull << 20;
    std::vector<abnn_synapse> buf(std::min<u64>(chunk, std::max<u64>(h->n_local, 1)));
    for (u64 off = 0; ok && off < h->n_local; off += chunk) {
        const u64 m = std::min(chunk, h->n_local - off);
        if (cudaMemcpyAsync(buf.data(), h->d_syn + off, m * sizeof(abnn_synapse), cudaMemcpyDeviceToHost, h->st) != cudaSuccess ||
            cudaStreamSynchronize(h->st) != cudaSuccess) { std::fclose(f); return fail(ABNN_ERR_CUDA, "device read failed during save"); }
        ok = std::fwrite(buf.data(), sizeof(abnn_synapse), m, f) == m;
    }
    ok = (std::fclose(f) == 0) && ok;
    return ok ? 0 : fail(ABNN_ERR_IO, std::string("short write: ") + path);
}

int abnn_load_bnn(abnn_handle* h, const char* path)
{
    RET(use(h));
    if (!path) return fail(ABNN_ERR_INVALID, "null path");
    if (h->p.world_size != 1) return fail(ABNN_ERR_UNSUPPORTED, ".bnn v1 holds one unsharded table; load into a single-GPU handle");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(ABNN_ERR_IO, std::string("cannot open: ") + path);
    uint32_t hdr[2] = {0, 0};
    if (std::fread(hdr, 4, 2, f) != 2) { std::fclose(f); return fail(ABNN_ERR_IO, "short header"); }
    // brain.cpp:174 rejects a file whose counts differ from the compiled-in shape. The neuron count must match here too;
    // the record count may be anything the handle can hold (abnn_save_bnn writes the LIVE count, which pruning and
    // synaptogenesis move away from n_syn).
    if (hdr[1] != h->N) {
        std::fclose(f);
        return fail(ABNN_ERR_SHAPE, ".bnn header N_NRN does not match the handle's shape");
    }
    const u64 n = hdr[0], chunk = 4ull << 20;
    if (n > h->cap) { std::fclose(f); return fail(ABNN_ERR_SHAPE, ".bnn header N_SYN exceeds the handle's synapse capacity"); }
    {   // the whole table must be there before the device table is touched
        const long at = std::ftell(f);
        std::fseek(f, 0, SEEK_END);
        const long end = std::ftell(f);
        std::fseek(f, at, SEEK_SET);
        if (at < 0 || end < 0 || (u64)(end - at) < n * sizeof(abnn_synapse)) { std::fclose(f); return fail(ABNN_ERR_IO, "short file: fewer records than the header announces"); }
    }
    std::vector<abnn_synapse> buf(std::min<u64>(chunk, std::max<u64>(n, 1)));
    for (u64 off = 0; off < n; off += chunk) {
        const u64 m = std::min(chunk, n - off);
        if (std::fread(buf.data(), sizeof(abnn_synapse), m, f) != m) { std::fclose(f); return fail(ABNN_ERR_IO, "short read"); }
        if (cudaMemcpyAsync(h->d_syn + off, buf.data(), m * sizeof(abnn_synapse), cudaMemcpyHostToDevice, h->st) != cudaSuccess ||
            cudaStreamSynchronize(h->st) != cudaSuccess) { std::fclose(f); return fail(ABNN_ERR_CUDA, "device write failed during load"); }
    }
    std::fclose(f);
    h->n_local = n; h->n_local_all.assign(1, n); h->counts_dirty = false;
    RET(check_table(h, false));
    RET(sort_table(h));
    RET(reset_structural(h));
    return 0;
}

// ---- .bnn v2: exact resume ------------------------------------------------------------------------
namespace {
struct Bnn2Header {
    char     magic[4];             // "BNN2"
    uint32_t version, scalars_size, params_size;
    uint64_t n_local, n_neuron, n_syn_global;
    uint32_t rank, world, n_output, fir_size;
    uint32_t grow_count, snapshot;
    uint32_t prune_count, pad_;
};
bool put(FILE* f, const void* p, size_t n) { return n == 0 || std::fwrite(p, 1, n, f) == n; }
bool get(FILE* f, void* p, size_t n) { return n == 0 || std::fread(p, 1, n, f) == n; }
// device -> file / file -> device through a bounded host buffer
int dev_to_file(abnn_handle* h, FILE* f, const void* d, size_t bytes, std::vector<char>& buf)
{
    for (size_t off = 0; off < bytes; off += buf.size()) {
        const size_t m = std::min(buf.size(), bytes - off);
        CU(cudaMemcpyAsync(buf.data(), (const char*)d + off, m, cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        if (!put(f, buf.data(), m)) return fail(ABNN_ERR_IO, "short write");
    }
    return 0;
}
int file_to_dev(abnn_handle* h, FILE* f, void* d, size_t bytes, std::vector<char>& buf)
{
    for (size_t off = 0; off < bytes; off += buf.size()) {
        const size_t m = std::min(buf.size(), bytes - off);
        if (!get(f, buf.data(), m)) return fail(ABNN_ERR_IO, "short read");
        CU(cudaMemcpyAsync((char*)d + off, buf.data(), m, cudaMemcpyHostToDevice, h->st));
        CU(cudaStreamSynchronize(h->st));
    }
    return 0;
}
}  // namespace

int abnn_save_state(abnn_handle* h, const char* path)
{
    RET(use(h));
    if (!path) return fail(ABNN_ERR_INVALID, "null path");
    RET(refresh_counts(h));
    RET(ensure_view(h));
    DevScalars sc; RET(read_scalars(h, &sc));
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(ABNN_ERR_IO, std::string("cannot open for writing: ") + path);
    Bnn2Header hd{};
    std::memcpy(hd.magic, "BNN2", 4);
    hd.version = 2; hd.scalars_size = sizeof(DevScalars); hd.params_size = sizeof(abnn_params);
    hd.n_local = h->n_local; hd.n_neuron = h->N;
    for (u64 v : h->n_local_all) hd.n_syn_global += v;
    hd.rank = h->p.rank; hd.world = h->p.world_size; hd.n_output = h->p.n_output; hd.fir_size = h->p.fir_size;
    hd.grow_count = h->d.grow ? std::min(sc.grow_count, h->grow_cap) : 0;
    hd.prune_count = h->d.prune_list ? std::min(sc.prune_count, PRUNE_CAP_DEFAULT) : 0;
    hd.snapshot = h->d.view != h->d.live;
    std::vector<char> buf(64u << 20);
    int rc = 0;
    bool ok = put(f, &hd, sizeof hd) && put(f, &h->p, sizeof h->p) && put(f, &sc, sizeof sc) &&
              put(f, h->n_local_all.data(), h->n_local_all.size() * sizeof(u64));
    if (!ok) rc = fail(ABNN_ERR_IO, "short write");
    const size_t no = h->p.n_output;
    if (!rc) rc = dev_to_file(h, f, h->rs.rate, no * sizeof(float), buf);
    if (!rc) rc = dev_to_file(h, f, h->rs.iir, no * sizeof(float), buf);
    if (!rc) rc = dev_to_file(h, f, h->rs.fir, (size_t)h->p.fir_size * no * sizeof(float), buf);
    if (!rc) rc = dev_to_file(h, f, h->rs.smooth, no * sizeof(float), buf);
    if (!rc) rc = dev_to_file(h, f, h->d.live, h->N * sizeof(u64), buf);
    if (!rc) rc = dev_to_file(h, f, h->d.visited, h->N * sizeof(u64), buf);
    if (!rc && hd.snapshot) rc = dev_to_file(h, f, h->d.view, h->N * sizeof(u64), buf);
    if (!rc && hd.grow_count) rc = dev_to_file(h, f, h->d.grow, (size_t)hd.grow_count * sizeof(GrowCand), buf);
    if (!rc && hd.prune_count) rc = dev_to_file(h, f, h->d.prune_list, (size_t)hd.prune_count * sizeof(u64), buf);
    if (!rc) rc = dev_to_file(h, f, h->d_syn, h->n_local * sizeof(abnn_synapse), buf);
    if (std::fclose(f) != 0 && !rc) rc = fail(ABNN_ERR_IO, std::string("short write: ") + path);
    return rc;
}

int abnn_load_state(abnn_handle* h, const char* path)
{
    RET(use(h));
    if (!path) return fail(ABNN_ERR_INVALID, "null path");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(ABNN_ERR_IO, std::string("cannot open: ") + path);
    Bnn2Header hd{};
    abnn_params saved{};
    DevScalars sc{};
    int rc = 0;
    if (!get(f, &hd, sizeof hd) || std::memcmp(hd.magic, "BNN2", 4) != 0 || hd.version != 2) rc = fail(ABNN_ERR_IO, "not a .bnn v2 file");
    else if (hd.scalars_size != sizeof(DevScalars) || hd.params_size != sizeof(abnn_params)) rc = fail(ABNN_ERR_SHAPE, ".bnn v2 written by a different library build");
    else if (hd.n_neuron != h->N || hd.rank != h->p.rank || hd.world != h->p.world_size || hd.n_output != h->p.n_output ||
             hd.fir_size != h->p.fir_size || hd.snapshot != (u32)(h->d.view != h->d.live))
        rc = fail(ABNN_ERR_SHAPE, ".bnn v2 header does not match the handle (neurons / rank / world / read-out / src view)");
    else if (hd.n_local > h->cap) rc = fail(ABNN_ERR_CAPACITY, "synapse table capacity exceeded");
    else if (hd.grow_count && (!h->d.grow || hd.grow_count > h->grow_buf)) rc = fail(ABNN_ERR_SHAPE, ".bnn v2 holds staged growth but the handle has p_new = 0");
    else if (hd.prune_count && (!h->d.prune_list || hd.prune_count > PRUNE_CAP_DEFAULT)) rc = fail(ABNN_ERR_SHAPE, ".bnn v2 holds staged prune candidates but the handle has compact_every <= 1");
    std::vector<u64> counts(h->p.world_size);
    if (!rc && !(get(f, &saved, sizeof saved) && get(f, &sc, sizeof sc) && get(f, counts.data(), counts.size() * sizeof(u64))))
        rc = fail(ABNN_ERR_IO, "short read");
    // The fields that define WHAT the saved state means must equal the handle's: the table is stored in table order (a
    // GIVEN-order table in a sorted handle would break the sorted insertion of later growth steps), the event stream
    // continues from the saved counters under the same sampler and seed, timestamps are in the saved clock's ticks.
    // Execution mode, learning rates and placement may differ (a PARALLEL run can be resumed in EXACT mode).
    if (!rc) {
        const abnn_params& a = saved; const abnn_params& b = h->p;
        if (a.sampler != b.sampler || a.release_rng != b.release_rng || a.clock_mode != b.clock_mode || a.src_view != b.src_view ||
            a.rbar_mode != b.rbar_mode || a.sample_block != b.sample_block || a.table_order != b.table_order || a.seed != b.seed ||
            a.window_pre != b.window_pre || a.refractory != b.refractory || a.teacher_gap != b.teacher_gap ||
            a.max_spikes_per_pass != b.max_spikes_per_pass || a.track_visits != b.track_visits || a.n_input != b.n_input ||
            a.n_hidden != b.n_hidden || (a.compact_every > 1 ? a.compact_every : 1) != (b.compact_every > 1 ? b.compact_every : 1))
            rc = fail(ABNN_ERR_SHAPE, ".bnn v2 was written under different semantics (sampler / sample_block / table_order / clock / src view / "
                                      "r-bar mode / seed / window / refractory / budget / visits / shape differ from the handle's)");
    }
    if (!rc) {   // everything the rest of the file must hold, checked before any device state is overwritten
        const size_t no_ = h->p.n_output;
        const u64 need = (u64)(3 + h->p.fir_size) * no_ * sizeof(float) + (u64)(hd.snapshot ? 3 : 2) * h->N * sizeof(u64) +
                         (u64)hd.grow_count * sizeof(GrowCand) + (u64)hd.prune_count * sizeof(u64) + hd.n_local * sizeof(abnn_synapse);
        const long at = std::ftell(f);
        std::fseek(f, 0, SEEK_END);
        const long end = std::ftell(f);
        std::fseek(f, at, SEEK_SET);
        if (at < 0 || end < 0 || (u64)(end - at) < need) rc = fail(ABNN_ERR_IO, ".bnn v2 file is shorter than its header announces");
    }
    std::vector<char> buf(64u << 20);
    const size_t no = h->p.n_output;
    if (!rc) rc = file_to_dev(h, f, h->rs.rate, no * sizeof(float), buf);
    if (!rc) rc = file_to_dev(h, f, h->rs.iir, no * sizeof(float), buf);
    if (!rc) rc = file_to_dev(h, f, h->rs.fir, (size_t)h->p.fir_size * no * sizeof(float), buf);
    if (!rc) rc = file_to_dev(h, f, h->rs.smooth, no * sizeof(float), buf);
    if (!rc) rc = file_to_dev(h, f, h->d.live, h->N * sizeof(u64), buf);
    if (!rc) rc = file_to_dev(h, f, h->d.visited, h->N * sizeof(u64), buf);
    if (!rc && hd.snapshot) rc = file_to_dev(h, f, h->d.view, h->N * sizeof(u64), buf);
    if (!rc && hd.grow_count) rc = file_to_dev(h, f, h->d.grow, (size_t)hd.grow_count * sizeof(GrowCand), buf);
    if (!rc && hd.prune_count) rc = file_to_dev(h, f, h->d.prune_list, (size_t)hd.prune_count * sizeof(u64), buf);
    if (!rc) rc = file_to_dev(h, f, h->d_syn, hd.n_local * sizeof(abnn_synapse), buf);
    std::fclose(f);
    if (rc) return rc;
    if (cudaMemcpyAsync(h->d.sc, &sc, sizeof sc, cudaMemcpyHostToDevice, h->st) != cudaSuccess || cudaStreamSynchronize(h->st) != cudaSuccess)
        return fail(ABNN_ERR_CUDA, "device write failed during load");
    h->n_local = hd.n_local; h->n_local_all = counts; h->counts_dirty = false;
    h->struct_steps = sc.struct_steps; h->n_sorted = sc.n_sorted; h->n_dead = sc.n_dead;
    h->slack_ready = false; h->fire_ready = false; h->view_stale = false;
    if (h->step_exec) { cudaGraphExecDestroy(h->step_exec); h->step_exec = nullptr; }
    return 0;
}

// ---- manifest (host only) --------------------------------------------------------------------------
namespace {
enum FieldType { F_U32, F_U64, F_F32, F_F64, F_I32 };
struct Field { const char* name; FieldType type; size_t off; };
#define FLD(n, t) {#n, t, offsetof(abnn_params, n)}
const Field kFields[] = {
    FLD(n_input, F_U32), FLD(n_output, F_U32), FLD(n_hidden, F_U64), FLD(n_syn, F_U64), FLD(syn_capacity, F_U64), FLD(seed, F_U64),
    FLD(sampler, F_U32), FLD(release_rng, F_U32), FLD(clock_mode, F_U32), FLD(exec_mode, F_U32), FLD(src_view, F_U32),
    FLD(rbar_mode, F_U32), FLD(max_spikes_per_pass, F_U32), FLD(track_visits, F_U32), FLD(window_pre, F_U64),
    FLD(refractory, F_U64), FLD(teacher_gap, F_U64), FLD(base_scale, F_F32), FLD(a_ltp, F_F32), FLD(a_ltd, F_F32),
    FLD(w_min, F_F32), FLD(w_max, F_F32), FLD(eta_home, F_F32), FLD(target_rate_hz, F_F32), FLD(home_tick_hz, F_F32),
    FLD(eta_reward, F_F32), FLD(alpha_rbar, F_F32), FLD(w_prune, F_F32), FLD(p_new, F_F32), FLD(w_init, F_F32),
    FLD(rate_alpha, F_F32), FLD(peak_decay, F_F32), FLD(peak_init, F_F32), FLD(use_fir, F_U32), FLD(fir_size, F_U32),
    FLD(reward_window, F_U32), FLD(filter_tau, F_F64), FLD(dt_sec, F_F64), FLD(loss0, F_F64), FLD(device, F_I32),
    FLD(l2_persist, F_U32), FLD(sample_block, F_U32), FLD(table_order, F_U32), FLD(prune_in_place, F_U32), FLD(exchange, F_U32),
    FLD(compact_every, F_U32),
};
#undef FLD
std::string trim(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
bool parse_number(std::string v, double* out, u64* out_u, bool* is_int)
{
    if (v.size() >= 2 && (v.front() == '"' || v.front() == '\'') && v.back() == v.front()) v = v.substr(1, v.size() - 2);
    std::string t;
    for (char c : v) if (c != '_') t.push_back(c);
    if (t.empty()) return false;
    char* end = nullptr;
    errno = 0;
    const unsigned long long u = std::strtoull(t.c_str(), &end, 10);
    if (errno == 0 && end && *end == 0 && t[0] != '-') { *out_u = u; *out = (double)u; *is_int = true; return true; }
    end = nullptr;
    const double d = std::strtod(t.c_str(), &end);
    if (!end || *end != 0) return false;
    *out = d; *out_u = d < 0 ? 0 : (u64)d; *is_int = false;
    return true;
}
}  // namespace

int abnn_params_from_manifest(const char* path, abnn_params* p, uint64_t* steps_out, uint64_t* tau_ltd_out)
{
    if (!path || !p) return fail(ABNN_ERR_INVALID, "null argument");
    if (p->struct_size != sizeof(abnn_params)) return fail(ABNN_ERR_INVALID, "abnn_params must hold defaults (abnn_default_params) before a manifest is applied");
    FILE* f = std::fopen(path, "r");
    if (!f) return fail(ABNN_ERR_IO, std::string("cannot open: ") + path);
    char line[4096];
    bool have_neurons = false;
    u64 neurons = 0;
    int lineno = 0;
    while (std::fgets(line, sizeof line, f)) {
        ++lineno;
        std::string s(line);
        if (s.empty() || std::isspace((unsigned char)s[0]) || s[0] == '#' || s[0] == '-') continue;   // nested block / list / comment
        const size_t hash = s.find('#');
        if (hash != std::string::npos) s = s.substr(0, hash);
        const size_t colon = s.find(':');
        if (colon == std::string::npos) continue;
        const std::string key = trim(s.substr(0, colon)), val = trim(s.substr(colon + 1));
        if (val.empty()) continue;                         // start of a nested block
        double d = 0; u64 u = 0; bool is_int = false;
        const bool num = parse_number(val, &d, &u, &is_int);
        auto need_num = [&]() -> int {
            if (num) return 0;
            std::fclose(f);
            return fail(ABNN_ERR_INVALID, std::string(path) + ":" + std::to_string(lineno) + ": `" + key + "` needs a number, got `" + val + "`");
        };
        if (key == "neurons")        { RET(need_num()); neurons = u; have_neurons = true; }
        else if (key == "synapses")  { RET(need_num()); p->n_syn = u; }
        else if (key == "tau_LTP")   { RET(need_num()); p->window_pre = u; }
        else if (key == "tau_LTD")   { RET(need_num()); if (tau_ltd_out) *tau_ltd_out = u; }
        else if (key == "alpha_LTP") { RET(need_num()); p->a_ltp = (float)d; }
        else if (key == "alpha_LTD") { RET(need_num()); p->a_ltd = (float)d; }
        else if (key == "steps")     { RET(need_num()); if (steps_out) *steps_out = u; }
        else if (key == "rng_seed")  { RET(need_num()); p->seed = u; }
        else {
            for (const Field& fd : kFields) {
                if (key != fd.name) continue;
                RET(need_num());
                char* base = reinterpret_cast<char*>(p) + fd.off;
                switch (fd.type) {
                    case F_U32: *reinterpret_cast<uint32_t*>(base) = (uint32_t)u; break;
                    case F_U64: *reinterpret_cast<uint64_t*>(base) = u; break;
                    case F_I32: *reinterpret_cast<int32_t*>(base) = (int32_t)d; break;
                    case F_F32: *reinterpret_cast<float*>(base) = (float)d; break;
                    case F_F64: *reinterpret_cast<double*>(base) = d; break;
                }
                break;
            }
        }
    }
    std::fclose(f);
    if (have_neurons) {
        const u64 io = (u64)p->n_input + p->n_output;
        if (neurons < io) return fail(ABNN_ERR_INVALID, "manifest: `neurons` is smaller than n_input + n_output");
        p->n_hidden = neurons - io;
    }
    return 0;
}

// ---- per-pass operations --------------------------------------------------------------------------
int abnn_inject_inputs(abnn_handle* h, const float* v, uint32_t n, float hz)
{
    RET(use(h));
    if (!v || n != h->p.n_input) return fail(ABNN_ERR_INVALID, "inject_inputs: need n_input values (brain.cpp:75)");
    const float pTick = hz * 1000u * 1000000000ull;      // hz * kTickNS * NSEC_PER_SEC (brain.cpp:76)
    float* dv = nullptr;
    RET(stage_vec(h, v, n, &dv));
    const KParams kp = make_kparams(h, 0);
    CU(launch_inject(kp, h->d, dv, n, pTick, nullptr, h->st));
    return 0;
}

int abnn_teacher_force(abnn_handle* h, const float* expected, uint32_t n, float rate)
{
    RET(use(h));
    if (!expected || n != h->p.n_output) return fail(ABNN_ERR_INVALID, "teacher_force: need n_output values");
    float* dv = nullptr;
    RET(stage_vec(h, expected, n, &dv));
    const KParams kp = make_kparams(h, 0);
    CU(launch_teacher(kp, h->d, dv, n, rate, h->p.teacher_gap, nullptr, h->st));
    return 0;
}

int abnn_set_reward(abnn_handle* h, float reward)
{
    RET(use(h));
    k_set_reward<<<1, 1, 0, h->st>>>(h->d.sc, reward);
    CU(cudaGetLastError());
    return 0;
}
int abnn_get_reward(abnn_handle* h, float* reward, float* rbar)
{
    RET(use(h));
    DevScalars s; RET(read_scalars(h, &s));
    if (reward) *reward = s.reward;
    if (rbar) *rbar = s.rbar;
    return 0;
}

// Everything one pass enqueues on the handle's stream. Capture-safe for PARALLEL execution (no host
// synchronisation, no allocation): abnn_engine_step records it into a CUDA graph.
static int enqueue_pass(abnn_handle* h, KParams kp, cudaEvent_t before_traverse, cudaEvent_t after_traverse, bool head_refreshed = false)
{
    if (slack_mode(h, kp)) {
        kp.use_slack = 1;
        kp.use_line32 = line32_mode(h, kp) ? 1u : 0u;
        const u64 head = (u64)h->p.n_input + h->p.n_output;
        u64 s1 = h->N;                  // gate words [0, s1) to (re)build from the snapshot
        if (h->slack_ready) {           // the last pass left the gate words; inject / teacher forcing touched the head since
            s1 = head_refreshed ? 0 : head;
        } else RET(ensure_view(h));
        if (kp.use_line32) {            // + fire32 / vis32 of the owned neurons: all of them, or the head after inject / teacher forcing
            u64 o0 = h->lo, o1 = h->hi;
            if (h->fire_ready) o1 = head_refreshed ? o0 : std::min(h->hi, std::max(h->lo, head));
            CU(launch_prepare32(kp, h->d, h->d.view, 0, s1, o0, o1, h->st));
        } else CU(launch_build_slack(kp, h->d, h->d.view, 0, s1, h->st));
    } else RET(ensure_view(h));
    if (before_traverse) CU(cudaEventRecord(before_traverse, h->st));
    switch (h->p.exec_mode) {
        case ABNN_EXEC_SERIAL:   CU(launch_traverse_serial(kp, h->d, h->st)); break;
        case ABNN_EXEC_PARALLEL: CU(launch_traverse_parallel(kp, h->d, h->sm_count, h->st)); break;
        default: RET(run_exact(h, kp)); break;
    }
    if (after_traverse) CU(cudaEventRecord(after_traverse, h->st));
    CU(launch_end_pass(kp, h->d.sc, h->d_stats, h->st));
    if (kp.use_line32) CU(launch_fold_prepare32(kp, h->d, h->lo, h->hi, h->st));   // fires / visits back into the 64-bit arrays + next pass's words
    RET(exchange_timestamps(h, kp));
    return 0;
}

int abnn_run_pass(abnn_handle* h, uint64_t events, abnn_pass_stats* stats)
{
    RET(use(h));
    RET(refresh_counts(h));
    const KParams kp = make_kparams(h, events);
    if (stats) CU(cudaEventRecord(h->ev0, h->st));
    RET(enqueue_pass(h, kp, stats ? h->evj : nullptr, stats ? h->evk : nullptr));
    if (stats) {
        CU(cudaEventRecord(h->ev1, h->st));
        CU(cudaMemcpyAsync(h->h_pin, h->d_stats, sizeof(abnn_pass_stats), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        std::memcpy(stats, h->h_pin, sizeof(abnn_pass_stats));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        stats->device_ms = ms;
        CU(cudaEventElapsedTime(&ms, h->evj, h->evk));
        stats->traverse_ms = ms;
    }
    return 0;
}

static ReadoutParams readout_params(const abnn_handle* h)
{
    ReadoutParams rp{};
    rp.n_input = h->p.n_input; rp.n_output = h->p.n_output;
    rp.rate_alpha = h->p.rate_alpha; rp.peak_decay = h->p.peak_decay;
    rp.use_fir = h->p.use_fir; rp.fir_size = h->p.fir_size; rp.reward_window = h->p.reward_window;
    rp.a = h->p.dt_sec / (h->p.filter_tau + h->p.dt_sec);                                      // rate-filter.h:29
    return rp;
}

// inject -> teacher forcing -> pass -> read-out step, all operands in device memory (d_frame)
static int enqueue_step(abnn_handle* h, const KParams& kp)
{
    const u32 ni = h->p.n_input, no = h->p.n_output;
    const bool refresh = slack_mode(h, kp) && h->slack_ready;
    const bool refresh_fire = refresh && h->fire_ready && line32_mode(h, kp);
    // the head's words are brought up to date by the prologue only if BOTH word sets are ready (else enqueue_pass rebuilds)
    CU(launch_step_prologue(kp, h->d, h->d_frame, ni, no, h->p.teacher_gap, refresh, refresh_fire, h->st));
    RET(enqueue_pass(h, kp, nullptr, nullptr, refresh));
    CU(launch_readout(kp, h->d, readout_params(h), h->rs, h->d_frame + ni, h->st));
    return 0;
}

int abnn_engine_step(abnn_handle* h, const float* in, const float* expected, float hz, float teacher_rate, uint64_t events, float* rates)
{
    RET(use(h));
    if (!in || !expected) return fail(ABNN_ERR_INVALID, "engine_step: need an input and an expected frame");
    RET(refresh_counts(h));
    const u32 ni = h->p.n_input, no = h->p.n_output, nf = ni + no + 2;
    if (!h->d_frame) {
        CU(cudaMalloc(&h->d_frame, nf * sizeof(float)));
        CU(cudaMallocHost(&h->h_frame, (size_t)RING * nf * sizeof(float)));
        for (int i = 0; i < RING; ++i) CU(cudaEventCreateWithFlags(&h->frame_ev[i], cudaEventDisableTiming));
    }
    // stage the frame: pinned ring slot -> fixed device buffer (the graph's kernels read it there)
    const int slot = h->frame_pos;
    h->frame_pos = (h->frame_pos + 1) % RING;
    CU(cudaEventSynchronize(h->frame_ev[slot]));
    float* hf = h->h_frame + (size_t)slot * nf;
    std::memcpy(hf, in, ni * sizeof(float));
    std::memcpy(hf + ni, expected, no * sizeof(float));
    hf[ni + no] = hz * 1000u * 1000000000ull;            // pTick = hz * kTickNS * NSEC_PER_SEC (brain.cpp:76)
    hf[ni + no + 1] = teacher_rate;
    CU(cudaMemcpyAsync(h->d_frame, hf, nf * sizeof(float), cudaMemcpyHostToDevice, h->st));
    CU(cudaEventRecord(h->frame_ev[slot], h->st));

    const KParams kp = make_kparams(h, events);
    static const bool no_graph = tune_env("ABNN_NO_GRAPH") != nullptr;
    // Single-GPU handles, and sharded handles on the peer-memory exchange (params.exchange = ABNN_EXCHANGE_PEER) once the
    // gate words are being exchanged (slack_ready): that sequence holds no NCCL call. With the NCCL exchange inside the
    // captured sequence a 2-rank run hung in this environment (NCCL 2.28.9, driver 580); such handles enqueue eagerly.
    const bool pre[3] = {h->slack_ready, h->view_stale, h->fire_ready};
    const bool sharded_ok = h->p2p && pre[0] && slack_mode(h, kp) && h->slice >= (u64)h->p.n_input + h->p.n_output &&
                            !tune_env("ABNN_FULL_EXCHANGE");
    const bool exact_ok = h->p.exec_mode == ABNN_EXEC_EXACT && h->p.world_size == 1;
    if (exact_ok) RET(ensure_exact_scratch(h, kp));      // allocation must not happen inside a capture
    const bool capturable = ((h->p.exec_mode == ABNN_EXEC_PARALLEL && (h->p.world_size == 1 || sharded_ok)) || exact_ok) && !no_graph;
    const bool match = h->step_exec && h->step_events == events && h->step_counts == h->n_local_all &&
                       h->step_pre[0] == pre[0] && h->step_pre[1] == pre[1] && h->step_pre[2] == pre[2];
    // record a graph only for a state that has just repeated (same events / table sizes / word state as the previous call):
    // a run whose table changes every step (structural plasticity after every pass) enqueues eagerly instead of
    // re-capturing every time
    const bool repeated = h->step_calls > 0 && h->last_events == events && h->last_counts == h->n_local_all &&
                          h->last_pre[0] == pre[0] && h->last_pre[1] == pre[1] && h->last_pre[2] == pre[2];
    h->last_events = events; h->last_counts = h->n_local_all;
    h->last_pre[0] = pre[0]; h->last_pre[1] = pre[1]; h->last_pre[2] = pre[2];
    ++h->step_calls;
    if (capturable && match) {
        CU(cudaGraphLaunch(h->step_exec, h->st));
        h->slack_ready = h->step_post[0]; h->view_stale = h->step_post[1]; h->fire_ready = h->step_post[2];
        ++h->step_replays;
    } else if (capturable && repeated) {
        // second call onwards (the first one warms up lazily configured kernels): record this state's
        // sequence once, then replay it for as long as events / table size / exchange state repeat
        if (h->step_exec) { cudaGraphExecDestroy(h->step_exec); h->step_exec = nullptr; }
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_step(h, kp);
        const cudaError_t e = cudaStreamEndCapture(h->st, &g);
        if (rc != 0 || e != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            h->slack_ready = pre[0]; h->view_stale = pre[1]; h->fire_ready = pre[2];
            if (rc != 0) return rc;
            return fail(ABNN_ERR_CUDA, std::string("engine_step: stream capture failed: ") + cudaGetErrorString(e));
        }
        const cudaError_t ei = cudaGraphInstantiate(&h->step_exec, g, 0);
        cudaGraphDestroy(g);
        if (ei != cudaSuccess) {
            h->step_exec = nullptr; h->slack_ready = pre[0]; h->view_stale = pre[1]; h->fire_ready = pre[2];
            return fail(ABNN_ERR_CUDA, "engine_step: cudaGraphInstantiate failed");
        }
        h->step_events = events; h->step_counts = h->n_local_all;
        h->step_pre[0] = pre[0]; h->step_pre[1] = pre[1]; h->step_pre[2] = pre[2];
        h->step_post[0] = h->slack_ready; h->step_post[1] = h->view_stale; h->step_post[2] = h->fire_ready;
        CU(cudaGraphLaunch(h->step_exec, h->st));
        ++h->step_replays;
    } else {
        RET(enqueue_step(h, kp));
    }
    if (rates) {
        CU(cudaMemcpyAsync(h->h_pin, h->rs.smooth, no * sizeof(float), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        std::memcpy(rates, h->h_pin, no * sizeof(float));
    }
    return 0;
}

int abnn_sync(abnn_handle* h)
{
    RET(use(h));
    CU(cudaStreamSynchronize(h->st));
    RET(p2p_check(h));
    return 0;
}

int abnn_timer_mark(abnn_handle* h, uint32_t slot)
{
    RET(use(h));
    if (slot >= 8) return fail(ABNN_ERR_INVALID, "timer slot out of range");
    CU(cudaEventRecord(h->timer[slot], h->st));
    return 0;
}
int abnn_timer_elapsed(abnn_handle* h, uint32_t a, uint32_t b, double* ms)
{
    RET(use(h));
    if (a >= 8 || b >= 8 || !ms) return fail(ABNN_ERR_INVALID, "bad timer arguments");
    CU(cudaEventSynchronize(h->timer[b]));
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, h->timer[a], h->timer[b]));
    *ms = f;
    return 0;
}

int abnn_read_outputs(abnn_handle* h, uint8_t* spikes, uint32_t n)
{
    RET(use(h));
    if (!spikes || n != h->p.n_output) return fail(ABNN_ERR_INVALID, "read_outputs: need n_output slots");
    const KParams kp = make_kparams(h, 0);
    CU(launch_read_outputs(kp, h->d, h->rs.spikes, n, h->st));
    CU(cudaMemcpyAsync(h->h_pin, h->rs.spikes, n, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    std::memcpy(spikes, h->h_pin, n);
    return 0;
}

static int readout_enqueue(abnn_handle* h, const float* expected, uint32_t n)
{
    if (n != h->p.n_output) return fail(ABNN_ERR_INVALID, "readout: need n_output values");
    float* de = nullptr;
    if (expected) RET(stage_vec(h, expected, n, &de));
    const KParams kp = make_kparams(h, 0);
    CU(launch_readout(kp, h->d, readout_params(h), h->rs, de, h->st));
    return 0;
}
int abnn_readout_step(abnn_handle* h, const float* expected, uint32_t n)
{
    RET(use(h));
    return readout_enqueue(h, expected, n);
}
int abnn_readout_filtered(abnn_handle* h, const float* expected, float* rates, uint32_t n)
{
    RET(use(h));
    RET(readout_enqueue(h, expected, n));
    CU(cudaMemcpyAsync(h->h_pin, h->rs.smooth, n * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    if (rates) std::memcpy(rates, h->h_pin, n * sizeof(float));
    return 0;
}
int abnn_get_loss(abnn_handle* h, double* last_loss, uint64_t* windows_done)
{
    RET(use(h));
    DevScalars s; RET(read_scalars(h, &s));
    if (last_loss) *last_loss = s.last_loss;
    if (windows_done) *windows_done = s.windows_done;
    return 0;
}

// ---- structural plasticity ------------------------------------------------------------------------
namespace {
// Growth candidates staged since the last structural step, from every rank, sorted by the tick ordinal of the firing
// event; *owned = how many of them target this rank's neurons (they sort first). Collective when world_size > 1.
int gather_growth(abnn_handle* h, GrowCand** list_out, u32* owned_out, bool other_overflow = false, const DevScalars* have = nullptr)
{
    *list_out = h->d.grow; *owned_out = 0;
    DevScalars sc;
    if (have) sc = *have; else RET(read_scalars(h, &sc));
    bool overflow = sc.grow_overflow != 0 || other_overflow;      // (the prune staging buffer of compact_every > 1 shares the agreement)
    u32 n = std::min(sc.grow_count, h->grow_cap);
    GrowCand* list = h->d.grow;
    u32 total = n;
    std::vector<u64> cnt(h->p.world_size);
    if (h->p.world_size > 1) {
        if (!h->comm) return fail(ABNN_ERR_COMM, "world_size > 1 but abnn_comm_init has not been called");
        // every rank needs everyone's candidates: exchange counts — with this rank's overflow flag in the top bit, so that
        // the ranks AGREE on the error before any of them leaves the collective sequence — pad to the maximum, allgather
        const u64 mine = (u64)n | (overflow ? 1ull << 63 : 0);
        CU(cudaMemcpyAsync(h->d_counts + h->p.rank, &mine, sizeof(u64), cudaMemcpyHostToDevice, h->st));
        NC(ncclAllGather(h->d_counts + h->p.rank, h->d_counts, 1, ncclUint64, h->comm, h->st));
        CU(cudaMemcpyAsync(cnt.data(), h->d_counts, sizeof(u64) * h->p.world_size, cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        for (u64& c : cnt) { if (c >> 63) overflow = true; c &= ~(1ull << 63); }
    }
    if (overflow) {                                          // the same status on every rank; the staged candidates are dropped
        k_reset_grow<<<1, 1, 0, h->st>>>(h->d.sc);
        k_set_struct<<<1, 1, 0, h->st>>>(h->d.sc, h->struct_steps, h->n_sorted, h->n_dead);    // drops the staged prune candidates too
        CU(cudaStreamSynchronize(h->st));
        return fail(ABNN_ERR_CAPACITY, "growth / prune staging buffer overflowed on a rank (candidates of this interval are dropped); call abnn_prune_and_grow more often");
    }
    if (h->p.world_size > 1) {
        u32 maxc = 0;
        for (u64 c : cnt) maxc = std::max<u32>(maxc, (u32)c);
        total = maxc * h->p.world_size;
        const u32 need = next_pow2(std::max<u32>(total, 1));
        if (need > h->grow_all_buf) {
            if (h->d_grow_all) CU(cudaFree(h->d_grow_all));
            h->d_grow_all = nullptr;
            CU(cudaMalloc(&h->d_grow_all, (size_t)need * sizeof(GrowCand)));
            h->grow_all_buf = need;
        }
        if (maxc) {
            if (maxc > n) { k_pad_grow<<<(maxc - n + 255) / 256, 256, 0, h->st>>>(h->d.grow, n, maxc); CU(cudaGetLastError()); }
            NC(ncclAllGather(h->d.grow, h->d_grow_all, (size_t)maxc * sizeof(GrowCand), ncclUint8, h->comm, h->st));
        }
        list = h->d_grow_all;
    }
    if (total) {
        // sized by the capacity of the list, not by this step's count: growing the scratch means cudaFree + cudaMalloc next
        // to a 16 GB table, which costs more than the whole structural step (measured: 105 ms instead of 9 ms)
        RET(ensure_scratch(h, grow_sort_scratch_bytes(std::max<u32>(next_pow2(total), h->p.world_size > 1 ? h->grow_all_buf : h->grow_buf))));
        CU(cudaMemsetAsync(h->d_total + 1, 0, sizeof(u64), h->st));
        CU(launch_grow_sort_count(list, total, (u32)h->lo, (u32)h->hi, reinterpret_cast<u32*>(h->d_total + 1), h->d_scratch, h->st));
        if (h->p.world_size == 1) {
            *owned_out = total;                              // one shard owns every neuron: no count to wait for
        } else {
            u64 owned64 = 0;
            CU(cudaMemcpyAsync(&owned64, h->d_total + 1, sizeof(u64), cudaMemcpyDeviceToHost, h->st));
            CU(cudaStreamSynchronize(h->st));
            *owned_out = (u32)owned64;
        }
    }
    *list_out = list;
    return 0;
}
}  // namespace

// Stable prune-compaction of the whole table: into the spare table (count pass + scatter pass, no chained scan) when
// there is memory for one, else in place (k_compact). Dead records (compact_every > 1) carry weights below w_prune and go
// with the rest. *kept = surviving records; h->n_local is updated.
static int prune_compact(abnn_handle* h, u64* kept_out)
{
    const bool in_place_only = h->p.prune_in_place != 0;
    const bool two_tables = 2 * h->cap * sizeof(abnn_synapse) <= h->mem_total / 2;
    abnn_synapse* spare = nullptr;
    if (!in_place_only && (h->d_spare || two_tables)) {
        spare = h->d_spare; h->d_spare = nullptr;
        if (!spare && cudaMalloc(&spare, h->cap * sizeof(abnn_synapse)) != cudaSuccess) { cudaGetLastError(); spare = nullptr; }
    }
    CompactArgs a{};
    a.in = h->d_syn; a.out = spare ? spare : h->d_syn; a.n = h->n_local; a.pred = KEEP_NOT_PRUNED; a.w_prune = h->p.w_prune;
    a.out_cap = h->cap;
    u64 kept = 0;
    cudaError_t e = cudaSuccess;
    int rc = ensure_scratch(h, spare ? compact2_scratch_bytes(h->cap) : compact_scratch_bytes(h->cap));   // by capacity: never regrown
    if (!rc) {
        e = spare ? launch_compact_two_pass(a, h->d_scratch, h->d_total, h->st) : launch_compact(a, h->d_scratch, h->d_total, h->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&kept, h->d_total, sizeof(u64), cudaMemcpyDeviceToHost, h->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    }
    if (rc || e != cudaSuccess) {
        if (spare) h->d_spare = spare;               // the table itself is untouched when the copy went elsewhere
        if (rc) return rc;
        return fail(ABNN_ERR_CUDA, std::string("prune: ") + cudaGetErrorString(e));
    }
    if (spare) {                                     // the old table becomes the spare of the next structural step
        h->d_spare = h->d_syn;
        h->d_syn = spare; h->d.syn = spare;
        if (h->step_exec) { cudaGraphExecDestroy(h->step_exec); h->step_exec = nullptr; }   // the captured pass holds the old pointer
    }
    h->n_local = kept;
    *kept_out = kept;
    return 0;
}

// compact_every > 1 (README.md:122-124 "remove, compact periodically"). Every step: the staged prune candidates that are
// still below w_prune die in place, the owned growth candidates are appended behind the table in tick order — the step
// touches what changed, not the table. Steps 0, K, 2K, ... additionally rebuild the table: dead and pruned records leave
// (the whole table is swept then, so candidates are not needed), the tail is merged into the ordered region in the stable
// order the eager mode keeps (DST_SORTED: sorted insertion behind the existing records of each destination; INTERLEAVED:
// re-derived; AS_GIVEN: the tail already is where it belongs).
// One host synchronisation per step that does not rebuild (the count of records that died); the bookkeeping kernel that
// follows it is stream-ordered and nobody waits for it.
static int lazy_structural_step(abnn_handle* h, abnn_structural_stats* s, GrowCand* list, u32 owned, const DevScalars& sc, bool grow)
{
    const bool rebuild = h->struct_steps % h->p.compact_every == 0;
    const bool prune = h->p.w_prune > 0.f;
    // 1. pruned records die in place (not on a rebuild step: its sweep removes everything below w_prune)
    u64 marked = 0;
    if (prune && !rebuild) {
        CU(launch_mark_dead(h->d.prune_list, std::min(sc.prune_count, PRUNE_CAP_DEFAULT), h->d_syn, h->p.w_prune, h->d_total + 2, h->st));
        CU(cudaMemcpyAsync(h->h_pin, h->d_total + 2, sizeof(u64), cudaMemcpyDeviceToHost, h->st));
    }
    // 2. the grown synapses join the tail, in tick order, as many as fit
    const u32 m = (u32)std::min<u64>(owned, h->cap - h->n_local);
    if (m) CU(launch_grow_append(list, m, h->d_syn, h->n_local, h->p.w_init, h->st));
    if (grow) { k_reset_grow<<<1, 1, 0, h->st>>>(h->d.sc); CU(cudaGetLastError()); }
    if (prune && !rebuild) {
        CU(cudaStreamSynchronize(h->st));
        std::memcpy(&marked, h->h_pin, sizeof(u64));
    }
    const u64 slots = h->n_local + m;                // table slots before any rebuild
    h->n_local = slots;
    s->appended = m; s->dropped = owned - m;
    if (!rebuild) {
        s->pruned = marked;
        h->n_dead += marked;
    } else {
        const u64 alive_before = slots - h->n_dead;
        const u64 tail = slots - h->n_sorted;
        if (h->p.table_order == ABNN_TABLE_DST_SORTED && prune && tail && tail < (1ull << 31) && h->n_sorted) {
            // the tail's survivors become the "new records" of the fused prune + sorted-insertion sweep of the ordered region
            // the first rebuild (step 0) sees one step's growth, the later ones compact_every steps' worth
            RET(ensure_merge_scratch(h, tail, std::min<u64>(2 * tail * (h->struct_steps ? 1 : h->p.compact_every), h->cap - h->n_sorted)));
            CompactArgs a{};
            a.in = h->d_syn + h->n_sorted; a.out = h->mg_new; a.n = tail; a.pred = KEEP_NOT_PRUNED; a.w_prune = h->p.w_prune; a.out_cap = h->mg_cap;
            RET(ensure_scratch(h, std::max(compact_scratch_bytes(h->cap), compact2_scratch_bytes(h->cap))));
            CU(launch_compact(a, h->d_scratch, h->d_total, h->st));
            u64 m2 = 0;
            CU(cudaMemcpyAsync(&m2, h->d_total, sizeof(u64), cudaMemcpyDeviceToHost, h->st));
            CU(cudaStreamSynchronize(h->st));
            u64 kept = 0;
            if (m2) RET(merge_grown(h, nullptr, (u32)m2, &kept, h->n_sorted));       // sets n_local = kept + m2
            else { h->n_local = h->n_sorted; RET(prune_compact(h, &kept)); }
        } else {
            u64 kept = slots;
            if (prune) RET(prune_compact(h, &kept));
            const bool resort = h->p.table_order == ABNN_TABLE_DST_INTERLEAVED ? (tail != 0 || kept != slots)
                                                                                 : (h->p.table_order == ABNN_TABLE_DST_SORTED && tail != 0);
            if (resort) RET(sort_table(h));
        }
        s->pruned = alive_before - h->n_local;
        h->n_sorted = h->n_local;
        h->n_dead = 0;
    }
    h->struct_steps += 1;
    if (rebuild) h->n_sorted = h->n_local;
    k_set_struct<<<1, 1, 0, h->st>>>(h->d.sc, h->struct_steps, h->n_sorted, h->n_dead);
    CU(cudaGetLastError());
    return 0;
}

int abnn_prune_and_grow(abnn_handle* h, abnn_structural_stats* out)
{
    RET(use(h));
    abnn_structural_stats s{};
    s.n_before = h->n_local;
    const bool prune = h->p.w_prune > 0.f && h->n_local, grow = h->p.p_new > 0.f;
    static const bool resort = tune_env("ABNN_GROW_RESORT") != nullptr;      // measurements: full radix re-sort instead of the merge
    static const bool no_fuse = tune_env("ABNN_NO_FUSED_PRUNE") != nullptr;  // measurements: prune and merge as two passes
    const bool lazy = h->p.compact_every > 1;
    DevScalars sc{};
    if (lazy) RET(read_scalars(h, &sc));
    // growth candidates first (they do not depend on the table): their number decides how the table is rewritten
    GrowCand* list = nullptr;
    u32 owned = 0;
    if (grow) RET(gather_growth(h, &list, &owned, lazy && sc.prune_overflow != 0, lazy ? &sc : nullptr));
    else if (lazy && sc.prune_overflow) {
        k_set_struct<<<1, 1, 0, h->st>>>(h->d.sc, h->struct_steps, h->n_sorted, h->n_dead);
        return fail(ABNN_ERR_CAPACITY, "prune staging buffer overflowed; call abnn_prune_and_grow more often");
    }
    if (lazy) {
        RET(lazy_structural_step(h, &s, list, owned, sc, grow));
    } else if (prune && owned && h->p.table_order == ABNN_TABLE_DST_SORTED && !resort && !no_fuse && owned <= h->cap - h->n_local) {
        // 1+2 fused: one pass over the table removes the pruned records and opens the slots of the new ones
        u64 kept = 0;
        RET(merge_grown(h, list, owned, &kept));
        s.pruned = s.n_before - kept;
        s.appended = owned;
    } else {
        // 1. prune: stable compaction
        if (prune) {
            u64 kept = 0;
            const u64 before = h->n_local;
            RET(prune_compact(h, &kept));
            s.pruned = before - kept;
            if (s.pruned && h->p.table_order == ABNN_TABLE_DST_INTERLEAVED) RET(sort_table(h));   // the interleaved order is re-derived after every change
        }
        // 2. grow: candidates in event order, as many as fit
        if (owned) {
            const u64 room = h->cap - h->n_local;
            const u32 m = (u32)std::min<u64>(owned, room);
            if (m && h->p.table_order == ABNN_TABLE_DST_SORTED && !resort) {
                RET(merge_grown(h, list, m));                                        // sorted insert, no re-sort of the table
            } else {
                CU(launch_grow_append(list, m, h->d_syn, h->n_local, h->p.w_init, h->st));
                h->n_local += m;
                if (m) RET(sort_table(h));
            }
            s.appended = m; s.dropped = owned - m;
        }
    }
    if (!lazy) {
        if (grow) {
            k_reset_grow<<<1, 1, 0, h->st>>>(h->d.sc);
            CU(cudaGetLastError());
        }
        CU(cudaStreamSynchronize(h->st));
        h->n_sorted = h->n_local;
    }
    s.n_after = h->n_local;
    h->counts_dirty = true;
    if (h->p.world_size == 1) h->n_local_all.assign(1, h->n_local);
    if (out) *out = s;
    return 0;
}

// ---- raw state access -----------------------------------------------------------------------------
int abnn_download_timestamps(abnn_handle* h, uint64_t* lf, uint64_t* lv)
{
    RET(use(h));
    RET(ensure_view(h));                 // collective when world_size > 1: every rank downloads together
    // lastFired: the replicated view is what every rank agrees on after a pass; single GPU: live.
    const u64* src = h->p.world_size > 1 ? h->d.view : h->d.live;
    if (lf) CU(cudaMemcpyAsync(lf, src, h->N * sizeof(u64), cudaMemcpyDeviceToHost, h->st));
    if (lv) CU(cudaMemcpyAsync(lv, h->d.visited, h->N * sizeof(u64), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}
int abnn_upload_timestamps(abnn_handle* h, const uint64_t* lf, const uint64_t* lv)
{
    RET(use(h));
    h->slack_ready = false; h->fire_ready = false;
    if (lf) h->view_stale = false;
    if (lf) {
        CU(cudaMemcpyAsync(h->d.live, lf, h->N * sizeof(u64), cudaMemcpyHostToDevice, h->st));
        if (h->d.view != h->d.live) CU(cudaMemcpyAsync(h->d.view, lf, h->N * sizeof(u64), cudaMemcpyHostToDevice, h->st));
    }
    if (lv) CU(cudaMemcpyAsync(h->d.visited, lv, h->N * sizeof(u64), cudaMemcpyHostToDevice, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}
int abnn_download_gate_words(abnn_handle* h, uint32_t* words, uint32_t* valid_out)
{
    RET(use(h));
    if (valid_out) *valid_out = h->slack_ready ? 1u : 0u;
    if (!words) return 0;
    if (!h->d.slack) return fail(ABNN_ERR_UNSUPPORTED, "this handle has no gate words (LIVE src view)");
    CU(cudaMemcpyAsync(words, h->d.slack, h->N * sizeof(u32), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}
int abnn_get_clock(abnn_handle* h, uint64_t* clock)
{
    RET(use(h));
    if (!clock) return fail(ABNN_ERR_INVALID, "null output");
    DevScalars s; RET(read_scalars(h, &s));
    *clock = s.clock;
    return 0;
}
int abnn_set_clock(abnn_handle* h, uint64_t clock)
{
    RET(use(h));
    h->slack_ready = false; h->fire_ready = false;   // the gate and fire words are relative to the clock
    k_set_clock<<<1, 1, 0, h->st>>>(h->d.sc, clock);
    CU(cudaGetLastError());
    return 0;
}

}  // extern "C"
