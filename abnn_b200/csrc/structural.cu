// abnn_b200/csrc/structural.cu — structural plasticity (README.md:120-127; absent from the reference
// code) as deterministic, scan-based stream operations:
//   * k_compact      : STABLE stream compaction of the synapse table in ONE pass over HBM
//                      (16 B read per record + 16 B write per kept record, both coalesced), single-pass
//                      chained scan with decoupled look-back. Works in place: a tile publishes its count only
//                      after its records are in registers, and a tile's destination range never
//                      reaches past its own source range, so no unread record is overwritten.
//                      Used for pruning (keep !(w < w_prune)) and for the dst-owner filter of
//                      abnn_upload_synapses.
//   * k_grow_*       : growth candidates staged by firing events are ordered by the tick ordinal of
//                      the event that produced them (radix sort) and appended in that order.
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "kernels.h"

namespace abnn {

namespace {
constexpr int CT = 256;                 // threads per tile
constexpr int CI = 8;                   // records per thread (128 B contiguous per thread)
constexpr u64 TILE = (u64)CT * CI;
#define FLAG_AGG    (1ull << 62)
#define FLAG_PREFIX (2ull << 62)
#define VAL_MASK    ((1ull << 62) - 1)

__device__ __forceinline__ bool keep_record(const CompactArgs& a, const uint4& r)
{
    if (a.pred == KEEP_NOT_PRUNED) return !(__uint_as_float(r.z) < a.w_prune);
    return r.y >= a.dst_lo && r.y < a.dst_hi;
}
}  // namespace

// Measured and reverted (profiles/r1_notes.md §7): persistent CTAs with the next ticket requested early and 4 CTAs per SM
// (sweep of 9.3e8 records 24.7 ms instead of 11.6 ms), and a look-back window of 8 x 32 descriptors per round trip
// (16.7 ms): the single-pass chained scan below is the fastest of the three.
__global__ void __launch_bounds__(CT) k_compact(const CompactArgs a, u32* ticket, volatile u64* desc, u64* total)
{
    __shared__ u32 s_tile;
    __shared__ u32 s_warp[CT / 32];
    __shared__ u64 s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);     // tiles are taken in table order
    __syncthreads();
    const u32 tile = s_tile;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // a warp owns 32 * CI consecutive records; lane l loads records l, l + 32, ... (coalesced 512-byte rows).
    // Stable rank inside the warp: rows before mine (popc of their keep ballots) + kept lanes before me in my row.
    const u64 wbase = (u64)tile * TILE + (u64)warp * (32 * CI);
    const uint4* in = reinterpret_cast<const uint4*>(a.in);
    uint4 rec[CI];
    unsigned bal[CI];
    u32 warp_total = 0;
#pragma unroll
    for (int j = 0; j < CI; ++j) {
        const u64 idx = wbase + (u64)j * 32 + lane;
        bool k = false;
        if (idx < a.n) {
            rec[j] = in[idx];
            k = keep_record(a, rec[j]);
        }
        bal[j] = __ballot_sync(0xffffffffu, k);
        warp_total += __popc(bal[j]);
        if (a.drop_hist) {                                   // fused prune + merge: removed records per neuron
            // one atomic per destination and row: in a dst-sorted table the removed records of a row share one or two
            // destinations, and per-record atomics on one address serialise (measured: 71M removals took 250 ms)
            const bool drop = idx < a.n && !k;
            const unsigned dm = __ballot_sync(0xffffffffu, drop);
            if (drop) {
                const unsigned peers = __match_any_sync(dm, rec[j].y);
                if (lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&a.drop_hist[rec[j].y - a.shift_lo + 1], (u32)__popc(peers));
            }
        }
    }
    if (lane == 0) s_warp[warp] = warp_total;
    __syncthreads();
    u32 warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < CT / 32; ++w) { const u32 v = s_warp[w]; if (w < (int)warp) warp_off += v; block_total += v; }

    // chained scan across tiles: warp 0 looks back 32 predecessors at a time
    if (warp == 0) {
        u64 run = 0;
        if (tile == 0) {
            if (lane == 0) desc[0] = FLAG_PREFIX | block_total;
        } else {
            if (lane == 0) desc[tile] = FLAG_AGG | block_total;
            long long p = (long long)tile - 1;
            while (true) {
                const long long idx = p - lane;
                u64 v = FLAG_PREFIX;                                       // before the table: prefix 0
                if (idx >= 0) v = desc[idx];
                const unsigned inval = __ballot_sync(0xffffffffu, (v >> 62) == 0);
                const unsigned pref  = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                if (pref) {
                    const int fp = __ffs(pref) - 1;
                    const unsigned need = fp == 31 ? 0xffffffffu : ((2u << fp) - 1u);
                    if (inval & need) continue;
                    u64 part = lane <= (unsigned)fp ? (v & VAL_MASK) : 0;
                    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    run += part;
                    break;
                }
                if (inval) continue;
                u64 part = v & VAL_MASK;
                for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                run += part;
                p -= 32;
            }
            if (lane == 0) desc[tile] = FLAG_PREFIX | (run + block_total);
        }
        if (lane == 0) {
            s_prefix = run;
            if ((u64)(tile + 1) * TILE >= a.n) *total = run + block_total;  // last tile
        }
    }
    __syncthreads();
    uint4* out = reinterpret_cast<uint4*>(a.out);
    u64 pos = s_prefix + warp_off;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < CI; ++j) {
        if ((bal[j] >> lane) & 1u) {
            u64 p = pos + __popc(bal[j] & lt);
            if (a.shift) p += a.shift[rec[j].y - a.shift_lo];           // fused prune + merge: room for the new records in front
            if (p < a.out_cap) out[p] = rec[j];
        }
        pos += __popc(bal[j]);
    }
}

size_t compact_scratch_bytes(u64 n)
{
    const u64 tiles = (n + TILE - 1) / TILE;
    return 16 + (size_t)(tiles + 1) * sizeof(u64);
}

cudaError_t launch_compact(const CompactArgs& a, void* scratch, u64* d_total, cudaStream_t st)
{
    const u64 tiles = (a.n + TILE - 1) / TILE;
    cudaError_t e = cudaMemsetAsync(scratch, 0, compact_scratch_bytes(a.n), st);
    if (e != cudaSuccess) return e;
    if (tiles == 0) return cudaMemsetAsync(d_total, 0, sizeof(u64), st);
    u32* ticket = reinterpret_cast<u32*>(scratch);
    u64* desc = reinterpret_cast<u64*>(reinterpret_cast<char*>(scratch) + 16);
    k_compact<<<(unsigned)tiles, CT, 0, st>>>(a, ticket, desc, d_total);
    return cudaGetLastError();
}

// ---- out-of-place stable compaction without a chained scan -------------------------------------------------
// Count pass (16 B read per record) -> exclusive scan of the per-tile counts (cub, ~0.5M values at 1e9 records) -> scatter
// pass (16 B read + 16 B written per kept record). 48 bytes per record instead of the 32 of k_compact, but every tile is
// independent: no ticket, no look-back, no barrier wait on a polling warp (which is what holds k_compact at a third of
// the copy bandwidth, profiles/r1_notes.md §7). Out of place only (a tile may overwrite records that another tile has not
// read yet), so the in-place prune keeps k_compact. Same tiling and the same in-tile ranks as k_compact.
__global__ void __launch_bounds__(CT) k_count_kept(const CompactArgs a, u64* tile_cnt)
{
    __shared__ u32 s_warp[CT / 32];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 wbase = (u64)blockIdx.x * TILE + (u64)warp * (32 * CI);
    const uint4* in = reinterpret_cast<const uint4*>(a.in);
    uint4 rec[CI];
#pragma unroll
    for (int j = 0; j < CI; ++j) {
        const u64 idx = wbase + (u64)j * 32 + lane;
        if (idx < a.n) rec[j] = __ldcs(in + idx);
    }
    u32 kept = 0;
#pragma unroll
    for (int j = 0; j < CI; ++j) {
        const u64 idx = wbase + (u64)j * 32 + lane;
        const bool k = idx < a.n && keep_record(a, rec[j]);
        kept += __popc(__ballot_sync(0xffffffffu, k));
        if (a.drop_hist) {                                   // removed records per neuron, one atomic per destination and row
            const bool drop = idx < a.n && !k;
            const unsigned dm = __ballot_sync(0xffffffffu, drop);
            if (drop) {
                const unsigned peers = __match_any_sync(dm, rec[j].y);
                if (lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&a.drop_hist[rec[j].y - a.shift_lo + 1], (u32)__popc(peers));
            }
        }
    }
    if (lane == 0) s_warp[warp] = kept;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
#pragma unroll
        for (int w = 0; w < CT / 32; ++w) t += s_warp[w];
        tile_cnt[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(CT) k_scatter_kept(const CompactArgs a, const u64* __restrict__ tile_off, u64* total)
{
    __shared__ u32 s_warp[CT / 32];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 wbase = (u64)blockIdx.x * TILE + (u64)warp * (32 * CI);
    const uint4* in = reinterpret_cast<const uint4*>(a.in);
    uint4* out = reinterpret_cast<uint4*>(a.out);
    uint4 rec[CI];
    unsigned bal[CI];
#pragma unroll
    for (int j = 0; j < CI; ++j) {
        const u64 idx = wbase + (u64)j * 32 + lane;
        if (idx < a.n) rec[j] = __ldcs(in + idx);
    }
    u32 warp_total = 0;
#pragma unroll
    for (int j = 0; j < CI; ++j) {
        const u64 idx = wbase + (u64)j * 32 + lane;
        bal[j] = __ballot_sync(0xffffffffu, idx < a.n && keep_record(a, rec[j]));
        warp_total += __popc(bal[j]);
    }
    if (lane == 0) s_warp[warp] = warp_total;
    __syncthreads();
    u32 warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < CT / 32; ++w) { const u32 v = s_warp[w]; if (w < (int)warp) warp_off += v; block_total += v; }
    const u64 base = tile_off[blockIdx.x];
    if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1) *total = base + block_total;
    u64 pos = base + warp_off;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < CI; ++j) {
        if ((bal[j] >> lane) & 1u) {
            u64 p = pos + __popc(bal[j] & lt);
            if (a.shift) p += a.shift[rec[j].y - a.shift_lo];
            if (p < a.out_cap) __stcs(out + p, rec[j]);
        }
        pos += __popc(bal[j]);
    }
}
size_t compact2_scratch_bytes(u64 n)
{
    const u64 tiles = (n + TILE - 1) / TILE;
    size_t scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan, (u64*)nullptr, (u64*)nullptr, (long long)tiles);
    return 2 * (size_t)(tiles + 1) * sizeof(u64) + ((scan + 255) & ~(size_t)255) + 256;
}
cudaError_t launch_compact_two_pass(const CompactArgs& a, void* scratch, u64* d_total, cudaStream_t st)
{
    const u64 tiles = (a.n + TILE - 1) / TILE;
    if (tiles == 0) return cudaMemsetAsync(d_total, 0, sizeof(u64), st);
    u64* cnt = reinterpret_cast<u64*>(scratch);
    u64* off = cnt + tiles + 1;
    void* tmp = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(off + tiles + 1) + 255) & ~(uintptr_t)255);
    size_t scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan, cnt, off, (long long)tiles);
    k_count_kept<<<(unsigned)tiles, CT, 0, st>>>(a, cnt);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, scan, cnt, off, (long long)tiles, st);
    if (e != cudaSuccess) return e;
    k_scatter_kept<<<(unsigned)tiles, CT, 0, st>>>(a, off, d_total);
    return cudaGetLastError();
}

// ---- ABNN_TABLE_DST_SORTED: stable sort of the table by destination neuron -------------------------
// LSD radix sort (cub) on key = dst with the 16-byte record as the value: stable, so records that
// share a destination keep their table order. One-off cost at graph load / after a growth step:
// 2 x ceil(bits/8) streaming passes over the table.
__global__ void k_extract_dst(const abnn_synapse* syn, u64 n, u32* keys)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) keys[i] = syn[i].dst;
}
size_t sort_by_dst_temp_bytes(u64 n)
{
    size_t b = 0;
    cub::DoubleBuffer<u32> k(nullptr, nullptr);
    cub::DoubleBuffer<uint4> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, b, k, v, (long long)n, 0, 32);
    return b;
}
cudaError_t launch_sort_by_dst(abnn_synapse* syn, abnn_synapse* alt, u32* keys, u32* keys_alt, u64 n, int key_bits,
                               void* tmp, size_t tmp_bytes, bool* result_in_alt, cudaStream_t st)
{
    *result_in_alt = false;
    if (n < 2) return cudaSuccess;
    u64 blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_extract_dst<<<(unsigned)blocks, 256, 0, st>>>(syn, n, keys);
    cub::DoubleBuffer<u32> k(keys, keys_alt);
    cub::DoubleBuffer<uint4> v(reinterpret_cast<uint4*>(syn), reinterpret_cast<uint4*>(alt));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (long long)n, 0, key_bits, st);
    if (e != cudaSuccess) return e;
    *result_in_alt = v.Current() != reinterpret_cast<uint4*>(syn);
    return cudaGetLastError();
}

// ---- ABNN_TABLE_DST_INTERLEAVED: from the dst-sorted table `in` to the interleaved one (include/abnn.h) ----------------
// start[d - lo] = first record of destination d in the sorted table (start[span] = n): thread i closes the gap between the
// destinations of records i-1 and i.
__global__ void __launch_bounds__(256) k_run_starts(const abnn_synapse* __restrict__ in, u64 n, u32 lo, u32 span, u64* __restrict__ start)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (u64)gridDim.x * blockDim.x) {
        const u32 d1 = i < n ? in[i].dst - lo : span;                  // start[d] = i for every d in (d0, d1]
        const long long d0 = i ? (long long)(in[i - 1].dst - lo) : -1;
        for (long long dd = d0 + 1; dd <= (long long)d1; ++dd) start[dd] = i;
    }
}
// Record i of the sorted table (destination d, r-th of its run) lands behind the rows 0..r-1 of its group of 8 neurons and
// the row-r records of the group's lower destinations: base + sum_k min(c_k, r) + #{k < d & 7 : c_k > r}.
template <u32 G>
__global__ void __launch_bounds__(256) k_interleave(const abnn_synapse* __restrict__ in, abnn_synapse* __restrict__ out, u64 n, u32 lo,
                                                    u32 span, const u64* __restrict__ start)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const uint4 rec = __ldcs(reinterpret_cast<const uint4*>(in + i));
        const u32 d = rec.y, g0 = d & ~(G - 1u);                        // groups are aligned on the GLOBAL neuron id
        const u64 r = i - start[d - lo];
        u64 pos = 0, base = 0;
        bool have_base = false;
#pragma unroll
        for (u32 k = 0; k < G; ++k) {
            const u32 nid = g0 + k;
            if (nid < lo || nid - lo >= span) continue;
            const u64 b = start[nid - lo], c = start[nid - lo + 1] - b;
            if (!have_base) { base = b; have_base = true; }
            pos += c < r ? c : r;
            if (k < (d & (G - 1u)) && c > r) ++pos;
        }
        __stcs(reinterpret_cast<uint4*>(out + base + pos), rec);
    }
}
cudaError_t launch_interleave_by_dst(const abnn_synapse* in, abnn_synapse* out, u64 n, u32 lo, u32 span, u32 group, u64* start, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    u64 blocks = (n + 256) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_run_starts<<<(unsigned)blocks, 256, 0, st>>>(in, n, lo, span, start);
    if (group == 16) k_interleave<16><<<(unsigned)blocks, 256, 0, st>>>(in, out, n, lo, span, start);
    else k_interleave<8><<<(unsigned)blocks, 256, 0, st>>>(in, out, n, lo, span, start);
    return cudaGetLastError();
}

// ---- ABNN_TABLE_DST_SORTED: insert m new records into the sorted table of n records (stable: behind the
// existing records of their destination) without re-sorting it. With cnt_less[d] = number of new records
// whose dst < d (histogram + prefix sum over the neurons, new records sorted by (dst, order)):
//   existing record i  ->  out[i + cnt_less[dst_i]]                       (one streaming pass, 32 B per record)
//   new record j       ->  out[upper_bound(existing dst, dst_j) + j]      (m binary searches)
// Out of place (the caller swaps the tables). 10^9 records: ~11 ms instead of ~70 ms for the radix re-sort.
__global__ void k_new_hist(const abnn_synapse* nw, u32 m, u32 dst_lo, u32* cnt)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) atomicAdd(&cnt[nw[j].dst - dst_lo + 1], 1u);
}
__global__ void __launch_bounds__(256) k_merge_existing(const abnn_synapse* __restrict__ syn, u64 n, u32 dst_lo,
                                                        const u32* __restrict__ cnt_less, abnn_synapse* __restrict__ out)
{
    const uint4* in = reinterpret_cast<const uint4*>(syn);
    uint4* o = reinterpret_cast<uint4*>(out);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const uint4 r = __ldcs(in + i);
        __stcs(o + i + cnt_less[r.y - dst_lo], r);
    }
}
__global__ void k_merge_new(const abnn_synapse* __restrict__ syn, u64 n, const abnn_synapse* __restrict__ nw, u32 m,
                            abnn_synapse* __restrict__ out)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const abnn_synapse r = nw[j];
    u64 lo = 0, hi = n;                                  // first existing record with dst > r.dst
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (syn[mid].dst <= r.dst) lo = mid + 1; else hi = mid;
    }
    out[lo + j] = r;
}
// Fused prune + merge: the existing table `syn` is the one BEFORE pruning; pruned_le[h + 1] = removed records with
// dst <= dst_lo + h (inclusive scan of the histogram k_compact filled), so upper_bound(syn, dst) - pruned_le is the
// number of KEPT existing records up to and including that destination.
__global__ void k_merge_new_pruned(const abnn_synapse* __restrict__ syn, u64 n, const abnn_synapse* __restrict__ nw, u32 m,
                                   u32 dst_lo, const u32* __restrict__ pruned_le, abnn_synapse* __restrict__ out)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const abnn_synapse r = nw[j];
    u64 lo = 0, hi = n;                                  // first existing record with dst > r.dst
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (syn[mid].dst <= r.dst) lo = mid + 1; else hi = mid;
    }
    out[lo - pruned_le[r.dst - dst_lo + 1] + j] = r;
}
size_t merge_scan_temp_bytes(u64 n_slots)
{
    size_t b = 0;
    cub::DeviceScan::InclusiveSum(nullptr, b, (u32*)nullptr, (u32*)nullptr, (long long)n_slots);
    return b;
}
cudaError_t launch_merge_sorted(const abnn_synapse* syn, u64 n, const abnn_synapse* nw_sorted, u32 m, u32 dst_lo, u32 dst_span,
                                u32* cnt /* dst_span + 1, zeroed */, void* scan_tmp, size_t scan_tmp_bytes, abnn_synapse* out,
                                int sm_count, cudaStream_t st)
{
    if (!m) return cudaSuccess;
    k_new_hist<<<(m + 255) / 256, 256, 0, st>>>(nw_sorted, m, dst_lo, cnt);
    cudaError_t e = cub::DeviceScan::InclusiveSum(scan_tmp, scan_tmp_bytes, cnt, cnt, (long long)dst_span + 1, st);
    if (e != cudaSuccess) return e;
    if (n) {
        u64 blocks = (n + 255) / 256;
        if (blocks > (u64)sm_count * 16) blocks = (u64)sm_count * 16;
        k_merge_existing<<<(unsigned)blocks, 256, 0, st>>>(syn, n, dst_lo, cnt, out);
    }
    k_merge_new<<<(m + 255) / 256, 256, 0, st>>>(syn, n, nw_sorted, m, out);
    return cudaGetLastError();
}

// Prune and sorted insertion in ONE pass over the table (out of place): kept record with stable rank p goes to
// out[p + cnt_less[dst]], new record j to out[kept records with dst <= dst_j + j]. Same table, bit for bit, as
// launch_compact (in place) followed by launch_merge_sorted, with half the HBM traffic (16 B read + 16 B written per
// record instead of twice that). Kept count lands in *d_total.
cudaError_t launch_prune_merge_sorted(const abnn_synapse* syn, u64 n, float w_prune, const abnn_synapse* nw_sorted, u32 m, u32 dst_lo,
                                      u32 dst_span, u32* cnt /* dst_span + 1, zeroed */, u32* pruned /* dst_span + 2, zeroed */,
                                      void* scan_tmp, size_t scan_tmp_bytes, void* compact_scratch, u64* d_total,
                                      abnn_synapse* out, u64 out_cap, cudaStream_t st)
{
    k_new_hist<<<(m + 255) / 256, 256, 0, st>>>(nw_sorted, m, dst_lo, cnt);
    cudaError_t e = cub::DeviceScan::InclusiveSum(scan_tmp, scan_tmp_bytes, cnt, cnt, (long long)dst_span + 1, st);
    if (e != cudaSuccess) return e;
    CompactArgs a{};
    a.in = syn; a.out = out; a.n = n; a.out_cap = out_cap; a.pred = KEEP_NOT_PRUNED; a.w_prune = w_prune;
    a.shift = cnt; a.shift_lo = dst_lo; a.drop_hist = pruned;
    static const bool lookback = tune_env("ABNN_COMPACT_LOOKBACK") != nullptr;     // measurements: the single-pass chained scan
    e = lookback ? launch_compact(a, compact_scratch, d_total, st) : launch_compact_two_pass(a, compact_scratch, d_total, st);
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::InclusiveSum(scan_tmp, scan_tmp_bytes, pruned, pruned, (long long)dst_span + 2, st);
    if (e != cudaSuccess) return e;
    k_merge_new_pruned<<<(m + 255) / 256, 256, 0, st>>>(syn, n, nw_sorted, m, dst_lo, pruned, out);
    return cudaGetLastError();
}

// ---- growth -------------------------------------------------------------------------------------
__global__ void k_grow_prepare(GrowCand* c, u32 n, u32 n_pow2, u32 dst_lo, u32 dst_hi, u32* owned)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pow2) return;
    bool mine = false;
    if (i < n) mine = c[i].dst >= dst_lo && c[i].dst < dst_hi;
    if (!mine) c[i].order = ~0ull;                     // foreign / padding entries sink to the end
    const unsigned m = __ballot_sync(__activemask(), mine);
    if (mine && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicAdd(owned, (u32)__popc(m));
}
__global__ void k_grow_append(const GrowCand* c, u32 m, abnn_synapse* syn, u64 at, float w_init)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) syn[at + i] = abnn_synapse{c[i].src, c[i].dst, w_init, 0.f};
}

// Sort of the candidates by `order` (unique tick ordinals; foreign and padding entries carry ~0 and sink to the end):
// cub LSD radix sort on the 64-bit key with the 16-byte candidate as the value — 8 passes over a list of at most a few
// million entries. (The bitonic network this replaces needed log2(n)·(log2(n)+1)/2 launches: 190 for 2^19 candidates,
// 0.8 ms of launch latency in a 9 ms structural step.)
__global__ void k_grow_keys(const GrowCand* c, u32 n, u64* keys)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = c[i].order;
}
static size_t grow_sort_temp_bytes(u32 n)
{
    size_t b = 0;
    cub::DoubleBuffer<u64> k(nullptr, nullptr);
    cub::DoubleBuffer<uint4> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, b, k, v, (int)n, 0, 64);
    return b;
}
size_t grow_sort_scratch_bytes(u32 n)
{
    const size_t n_al = ((size_t)n + 31) & ~(size_t)31;
    return 2 * n_al * sizeof(u64) + n_al * sizeof(uint4) + grow_sort_temp_bytes(n) + 512;
}
cudaError_t launch_grow_sort_count(GrowCand* c, u32 n, u32 dst_lo, u32 dst_hi, u32* d_owned, void* scratch, cudaStream_t st)
{
    static_assert(sizeof(GrowCand) == sizeof(uint4), "candidates travel as 16-byte values");
    if (!n) return cudaSuccess;
    const unsigned blocks = (n + 255) / 256;
    const size_t n_al = ((size_t)n + 31) & ~(size_t)31;
    u64* keys = reinterpret_cast<u64*>(scratch);
    uint4* alt = reinterpret_cast<uint4*>(keys + 2 * n_al);
    void* tmp = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(alt + n_al) + 255) & ~(uintptr_t)255);
    k_grow_prepare<<<blocks, 256, 0, st>>>(c, n, n, dst_lo, dst_hi, d_owned);
    k_grow_keys<<<blocks, 256, 0, st>>>(c, n, keys);
    cub::DoubleBuffer<u64> k(keys, keys + n_al);
    cub::DoubleBuffer<uint4> v(reinterpret_cast<uint4*>(c), alt);
    size_t tmp_bytes = grow_sort_temp_bytes(n);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (int)n, 0, 64, st);
    if (e != cudaSuccess) return e;
    if (v.Current() != reinterpret_cast<uint4*>(c))
        e = cudaMemcpyAsync(c, v.Current(), (size_t)n * sizeof(uint4), cudaMemcpyDeviceToDevice, st);
    return e != cudaSuccess ? e : cudaGetLastError();
}
// compact_every > 1: the staged prune candidates (records written below w_prune since the last structural step) that are
// still below it become dead in place; *n_marked counts them (a record staged twice is marked once).
__global__ void k_mark_dead(const u64* __restrict__ list, u32 n, abnn_synapse* syn, float w_prune, u64* n_marked)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    abnn_synapse* r = syn + list[i];
    if (r->w < w_prune && atomicExch(&r->src, DEAD_SRC) != DEAD_SRC) atomicAdd(n_marked, 1ull);
}
cudaError_t launch_mark_dead(const u64* list, u32 n, abnn_synapse* syn, float w_prune, u64* n_marked, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(n_marked, 0, sizeof(u64), st);
    if (e != cudaSuccess || !n) return e;
    k_mark_dead<<<(n + 255) / 256, 256, 0, st>>>(list, n, syn, w_prune, n_marked);
    return cudaGetLastError();
}

// Every record of an uploaded / loaded table must name neurons of the handle: src < n_neuron (or the dead mark, where the
// caller allows it), lo <= dst < hi. *bad counts the records that do not (the kernels index per-neuron arrays with both).
__global__ void __launch_bounds__(256) k_validate_table(const abnn_synapse* __restrict__ syn, u64 n, u32 n_neuron, u32 lo, u32 hi,
                                                        u32 allow_dead, u64* bad)
{
    u32 mine = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const uint2 sd = *reinterpret_cast<const uint2*>(syn + i);
        const bool ok = (sd.x < n_neuron || (allow_dead && sd.x == DEAD_SRC)) && sd.y >= lo && sd.y < hi;
        mine += !ok;
    }
    if (mine) atomicAdd(reinterpret_cast<unsigned long long*>(bad), (unsigned long long)mine);
}
cudaError_t launch_validate_table(const abnn_synapse* syn, u64 n, u32 n_neuron, u32 lo, u32 hi, bool allow_dead, u64* bad, int sm_count,
                                  cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(u64), st);
    if (e != cudaSuccess || !n) return e;
    u64 blocks = (n + 255) / 256;
    if (blocks > (u64)sm_count * 8) blocks = (u64)sm_count * 8;
    k_validate_table<<<(unsigned)blocks, 256, 0, st>>>(syn, n, n_neuron, lo, hi, allow_dead ? 1u : 0u, bad);
    return cudaGetLastError();
}

cudaError_t launch_grow_append(const GrowCand* c, u32 m, abnn_synapse* syn, u64 at, float w_init, cudaStream_t st)
{
    if (!m) return cudaSuccess;
    k_grow_append<<<(m + 255) / 256, 256, 0, st>>>(c, m, syn, at, w_init);
    return cudaGetLastError();
}

}  // namespace abnn
