// abnn_b200/csrc/kernels.h — host-callable launchers of the CUDA kernels (internal to the library).
#pragma once
#include "common.cuh"

namespace abnn {

// traversal.cu
cudaError_t launch_traverse_parallel(const KParams& kp, const DevPtrs& d, int sm_count, cudaStream_t st);
cudaError_t launch_traverse_serial(const KParams& kp, const DevPtrs& d, cudaStream_t st);
// slack[n], n in [n0, n1), from lastFired values src[n] for the pass that starts at sc->clock (see k_build_slack)
cudaError_t launch_build_slack(const KParams& kp, const DevPtrs& d, const u64* src, u64 n0, u64 n1, cudaStream_t st);
bool line_kernel_selected(const KParams& kp);
// k_traverse_line32 (32-bit pass-relative timestamps): can it run this pass / conversions around the pass
bool line32_selected(const KParams& kp);
// gate words of [s0,s1) from the snapshot `src`, fire32 / vis32 of the owned neurons [o0,o1) from the 64-bit lastFired
cudaError_t launch_prepare32(const KParams& kp, const DevPtrs& d, const u64* src, u64 s0, u64 s1, u64 o0, u64 o1, cudaStream_t st);
// after k_end_pass: fires and visits of the pass back into lastFired / snapshot / lastVisited of the owned neurons [o0,o1),
// and their gate / fire / visit words for the next pass
cudaError_t launch_fold_prepare32(const KParams& kp, const DevPtrs& d, u64 o0, u64 o1, cudaStream_t st);
cudaError_t launch_end_pass(const KParams& kp, DevScalars* sc, abnn_pass_stats* out, cudaStream_t st);

// exchange.cu — the per-pass exchange of a sharded PARALLEL run as peer-memory stores (opt-in, see capi.cu:p2p_setup)
constexpr u32 P2P_MAX_WORLD = 8;
enum : u32 { P2P_DONE = 0, P2P_PUSHED = P2P_MAX_WORLD, P2P_EPOCH = 2 * P2P_MAX_WORLD, P2P_CTAS, P2P_ERROR, P2P_WORDS = 32 };   // u64 words of a rank's flag block
struct P2PTable {
    u32* slack[P2P_MAX_WORLD];      // every rank's gate-word array (own pointer at [rank], IPC-mapped peers elsewhere)
    u64* view[P2P_MAX_WORLD];       // every rank's 64-bit snapshot (the input/output head is delivered there)
    u64* flags[P2P_MAX_WORLD];      // every rank's flag block (P2P_WORDS u64)
    u32 world, rank;
};
// gate words of neurons [n0, n1) (the owned slice) into every rank's array, rank 0's lastFired[0, head) into every snapshot,
// fenced by two flag rounds; returns when everything is enqueued on st
cudaError_t launch_p2p_exchange(const KParams& kp, const DevPtrs& d, const P2PTable& t, u64 n0, u64 n1, u64 head, int sm_count,
                                cudaStream_t st);

// *bad = records whose src is not a neuron of the handle (or the dead mark, if allowed) or whose dst is outside [lo, hi)
cudaError_t launch_validate_table(const abnn_synapse* syn, u64 n, u32 n_neuron, u32 lo, u32 hi, bool allow_dead, u64* bad, int sm_count,
                                  cudaStream_t st);

// exact.cu — EXACT execution (two-phase, bit-identical to SERIAL)
size_t exact_scan_temp_bytes(u64 span);
// events still open against the pass-start lastFired into list[0 .. counter[0]) as (dst << 32 | event), counted per destination in
// cnt[dst - lo]; counter[1] = events that passed the pre-spike window
// (slot[j] = arrival number of list[j] at its destination)
cudaError_t launch_exact_phase1(const KParams& kp, const DevPtrs& d, u64* list, u32* slot, u32* cnt, u32 lo, u32 span, u32* counter,
                                int sm_count, cudaStream_t st);
// cursor = exclusive scan of cnt, then bucket[cursor[n] + slot] = the event indices grouped by destination
cudaError_t launch_exact_group(const u64* list, const u32* slot, const u32* counter, const u32* cnt, u32* cursor, u32 lo, u32 span,
                               u32* bucket, void* tmp, size_t tmp_bytes, int sm_count, cudaStream_t st);
cudaError_t launch_exact_phase3(const KParams& kp, const DevPtrs& d, u32* bucket, const u32* cnt, const u32* cursor, u32 lo, u32 span,
                                const u32* counter, int sm_count, cudaStream_t st);

// io_kernels.cu
struct ReadoutParams {
    u32 n_input, n_output;
    float rate_alpha, peak_decay;
    u32 use_fir, fir_size, reward_window;
    double a;                       // dt / (tau + dt)   (rate-filter.h:29)
};
struct ReadoutState { float* rate; float* iir; float* fir; float* smooth; unsigned char* spikes; };
// scal (nullable): read pTick / rate from device memory instead of the by-value argument (graph replay)
cudaError_t launch_inject(const KParams& kp, const DevPtrs& d, const float* v, u32 n, float pTick, const float* scal, cudaStream_t st);
cudaError_t launch_teacher(const KParams& kp, const DevPtrs& d, const float* expected, u32 n, float rate, u64 gap, const float* scal,
                           cudaStream_t st);
// inject + teacher forcing (+ gate-word refresh of the head) in one launch; frame = [in | expected | pTick | rate]
cudaError_t launch_step_prologue(const KParams& kp, const DevPtrs& d, const float* frame, u32 n_in, u32 n_out, u64 gap,
                                 bool refresh_slack, bool refresh_fire, cudaStream_t st);
cudaError_t launch_read_outputs(const KParams& kp, const DevPtrs& d, unsigned char* spikes, u32 n_out, cudaStream_t st);
cudaError_t launch_readout(const KParams& kp, const DevPtrs& d, const ReadoutParams& rp, const ReadoutState& rs,
                           const float* expected, cudaStream_t st);

// structural.cu
enum CompactPredicate : u32 { KEEP_NOT_PRUNED = 0, KEEP_OWNED = 1 };
struct CompactArgs {
    const abnn_synapse* in;   // may alias out (in-place, out <= in)
    abnn_synapse* out;
    u64 n;
    u64 out_cap;              // records `out` can hold; writes past it are dropped (total still counts them)
    u32 pred;
    float w_prune;
    u32 dst_lo, dst_hi;
    // fused prune + sorted merge (launch_prune_merge_sorted; out of place only): a kept record lands shift[dst - shift_lo]
    // slots behind its stable rank, a removed one is counted in drop_hist[dst - shift_lo + 1]. Null = plain compaction.
    const u32* shift;
    u32 shift_lo;
    u32* drop_hist;
};
size_t compact_scratch_bytes(u64 n);
// Stable stream compaction (single pass, decoupled look-back). Kept count lands in *d_total.
cudaError_t launch_compact(const CompactArgs& a, void* scratch, u64* d_total, cudaStream_t st);
// The same compaction out of place as count pass + tile-offset scan + scatter pass (no chained scan); `out` must not alias `in`.
size_t compact2_scratch_bytes(u64 n);
cudaError_t launch_compact_two_pass(const CompactArgs& a, void* scratch, u64* d_total, cudaStream_t st);
// Growth candidates c[0..n): entries whose dst is outside [dst_lo,dst_hi) get order = +inf, the owned ones are counted
// into *d_owned (zeroed by the caller) and the list is sorted by `order` (radix sort; scratch = grow_sort_scratch_bytes(n)).
size_t grow_sort_scratch_bytes(u32 n);
cudaError_t launch_grow_sort_count(GrowCand* c, u32 n, u32 dst_lo, u32 dst_hi, u32* d_owned, void* scratch, cudaStream_t st);
cudaError_t launch_grow_append(const GrowCand* c, u32 m, abnn_synapse* syn, u64 at, float w_init, cudaStream_t st);
// compact_every > 1: staged prune candidates list[0..n) that are still below w_prune become dead in place; *n_marked = how many
cudaError_t launch_mark_dead(const u64* list, u32 n, abnn_synapse* syn, float w_prune, u64* n_marked, cudaStream_t st);

// Stable sort of n records by dst (ABNN_TABLE_DST_SORTED). alt/keys/keys_alt: n-element scratch buffers.
// The sorted table ends up in `alt` when *result_in_alt (the caller copies it back), else in `syn`.
size_t sort_by_dst_temp_bytes(u64 n);
cudaError_t launch_sort_by_dst(abnn_synapse* syn, abnn_synapse* alt, u32* keys, u32* keys_alt, u64 n, int key_bits,
                               void* tmp, size_t tmp_bytes, bool* result_in_alt, cudaStream_t st);

// ABNN_TABLE_DST_INTERLEAVED: the dst-sorted table `in` (n records, destinations in [lo, lo + span)) interleaved per group of
// `group` (8 or 16) neurons into `out` (must not alias); start: span + 1 u64 of scratch (run starts).
cudaError_t launch_interleave_by_dst(const abnn_synapse* in, abnn_synapse* out, u64 n, u32 lo, u32 span, u32 group, u64* start, cudaStream_t st);

// Stable insertion of m new records (sorted by dst, ties in append order) into the dst-sorted table of n records,
// out of place: out[0 .. n+m). cnt: dst_span + 1 zeroed u32 slots (per-neuron histogram -> prefix sums).
size_t merge_scan_temp_bytes(u64 n_slots);
cudaError_t launch_merge_sorted(const abnn_synapse* syn, u64 n, const abnn_synapse* nw_sorted, u32 m, u32 dst_lo, u32 dst_span,
                                u32* cnt, void* scan_tmp, size_t scan_tmp_bytes, abnn_synapse* out, int sm_count, cudaStream_t st);

// Pruning (w < w_prune) and the sorted insertion above in one pass over the table; pruned: dst_span + 2 zeroed u32 slots,
// scan_tmp sized for dst_span + 2 slots, compact_scratch = compact_scratch_bytes(n). Kept count lands in *d_total.
cudaError_t launch_prune_merge_sorted(const abnn_synapse* syn, u64 n, float w_prune, const abnn_synapse* nw_sorted, u32 m, u32 dst_lo,
                                      u32 dst_span, u32* cnt, u32* pruned, void* scan_tmp, size_t scan_tmp_bytes,
                                      void* compact_scratch, u64* d_total, abnn_synapse* out, u64 out_cap, cudaStream_t st);

// init.cu
cudaError_t launch_init_er_beta(abnn_synapse* syn, u64 g0, u64 count, u64 seed, u64 n_neuron, u64 dlo, u64 dhi,
                                int sm_count, cudaStream_t st);

}  // namespace abnn
