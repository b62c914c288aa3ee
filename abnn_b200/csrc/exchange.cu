// abnn_b200/csrc/exchange.cu — the per-pass exchange of a sharded PARALLEL run as peer-memory stores over NVLink
// instead of an NCCL collective (abnn_params.exchange = ABNN_EXCHANGE_PEER, see capi.cu:p2p_setup).
//
// What the exchange has to deliver before the next pass starts (capi.cu:exchange_timestamps): every rank's gate-word
// array holds the words of EVERY neuron for the next pass, and every rank's snapshot holds rank 0's lastFired of the
// input/output head. Each rank owns a slice of the neurons; it converts its slice to gate words (the clock has already
// advanced) and stores them straight into all peers' arrays (IPC-mapped), then raises a flag at every peer.
//
// A rank that finishes its pass early must not overwrite words a slower peer's traversal kernel is still reading, so
// there are two flag rounds per exchange, both counted by an epoch that every rank increments once per exchange:
//   DONE   [r] at peer p = rank r has finished the traversal of this pass (nobody writes into p before all DONE)
//   PUSHED [r] at peer p = rank r's words have landed in p's arrays      (p starts its next pass after all PUSHED)
// Writers order data before flag with __threadfence_system(); readers poll with volatile loads (L1 bypass).
// Every spin is bounded (P2P_SPIN_LIMIT clock ticks, ~2 s): on expiry the kernel sets the error word and goes on, the
// host reports ABNN_ERR_COMM at the next synchronising call — a peer that never arrives must not hang the GPU.
#include "common.cuh"
#include "kernels.h"

namespace abnn {

namespace {
constexpr long long P2P_SPIN_LIMIT = 4000000000ll;      // clock64 ticks

__device__ __forceinline__ void wait_all(volatile u64* flags, u32 world, u64 epoch, u64* err)
{
    const long long t0 = clock64();
    for (u32 r = 0; r < world; ++r) {
        while (flags[r] < epoch) {
            if (clock64() - t0 > P2P_SPIN_LIMIT) { *err = 1; return; }
            __nanosleep(64);
        }
    }
    __threadfence_system();                                  // acquire: the peers' data stores are visible after their flags
}
}  // namespace

// One thread: this rank is done reading its gate words for the pass; bump the epoch and tell every peer.
__global__ void k_p2p_done(P2PTable t)
{
    if (blockIdx.x || threadIdx.x) return;
    u64* mine = t.flags[t.rank];
    const u64 epoch = mine[P2P_EPOCH] + 1;
    mine[P2P_EPOCH] = epoch;
    mine[P2P_CTAS] = 0;
    __threadfence_system();
    for (u32 r = 0; r < t.world; ++r) *(volatile u64*)(t.flags[r] + P2P_DONE + t.rank) = epoch;
}

// Convert the owned slice to gate words and store them into every rank's array; rank 0 also delivers the lastFired head.
// The last CTA to finish raises PUSHED at every peer.
__global__ void __launch_bounds__(256) k_p2p_push(const __grid_constant__ KParams kp, const DevPtrs d, P2PTable t, u64 n0, u64 n1, u64 head)
{
    __shared__ u32 s_last;
    u64* mine = t.flags[t.rank];
    const u64 epoch = *(volatile u64*)(mine + P2P_EPOCH);
    if (threadIdx.x == 0) wait_all(mine + P2P_DONE, t.world, epoch, mine + P2P_ERROR);
    __syncthreads();
    const u64 clock = d.sc->clock;                           // already advanced by k_end_pass: the next pass's first tick
    for (u64 n = n0 + (u64)blockIdx.x * blockDim.x + threadIdx.x; n < n1; n += (u64)gridDim.x * blockDim.x) {
        const u32 w = slack_word(clock, d.live[n], kp.window_pre);
        for (u32 r = 0; r < t.world; ++r) t.slack[r][n] = w;
    }
    if (t.rank == 0)
        for (u64 n = (u64)blockIdx.x * blockDim.x + threadIdx.x; n < head; n += (u64)gridDim.x * blockDim.x) {
            const u64 v = d.live[n];
            for (u32 r = 0; r < t.world; ++r) t.view[r][n] = v;
        }
    __threadfence_system();                                  // release: this thread's remote stores before the flag
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd((unsigned long long*)(mine + P2P_CTAS), 1ull) + 1 == gridDim.x;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence_system();
        for (u32 r = 0; r < t.world; ++r) *(volatile u64*)(t.flags[r] + P2P_PUSHED + t.rank) = epoch;
    }
}

// One thread: every rank's words have landed here; what follows on the stream may read the gate words.
__global__ void k_p2p_wait(P2PTable t)
{
    if (blockIdx.x || threadIdx.x) return;
    u64* mine = t.flags[t.rank];
    wait_all(mine + P2P_PUSHED, t.world, *(volatile u64*)(mine + P2P_EPOCH), mine + P2P_ERROR);
}

cudaError_t launch_p2p_exchange(const KParams& kp, const DevPtrs& d, const P2PTable& t, u64 n0, u64 n1, u64 head, int sm_count,
                                cudaStream_t st)
{
    k_p2p_done<<<1, 32, 0, st>>>(t);
    u64 blocks = (n1 - n0 + 255) / 256;
    if (blocks > (u64)sm_count) blocks = (u64)sm_count;      // all CTAs resident: the first thread of each spins on DONE
    if (blocks < 1) blocks = 1;
    k_p2p_push<<<(unsigned)blocks, 256, 0, st>>>(kp, d, t, n0, n1, head);
    k_p2p_wait<<<1, 32, 0, st>>>(t);
    return cudaGetLastError();
}

}  // namespace abnn
