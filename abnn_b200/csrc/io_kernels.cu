// abnn_b200/csrc/io_kernels.cu — the small per-pass kernels around the traversal: input injection,
// teacher forcing, output read-out and the rate/FIR filter + reward step. All O(n_input + n_output);
// they exist so that a pass needs no host access to device state (the reference pokes shared
// Metal buffers from the CPU every pass: brain.cpp:73-83,145-157, brain-engine.cpp:119-186).
#include "common.cuh"
#include "kernels.h"

namespace abnn {

// Brain::inject_inputs (brain.cpp:73-83): lf[i] = now iff u < pTick * v[i].
// `scal` (nullable): the scalar argument read from device memory instead, so that a captured graph can be
// replayed with a new value (abnn_engine_step).
__global__ void k_inject(const __grid_constant__ KParams kp, const DevPtrs d, const float* __restrict__ v, u32 n, float pTick,
                         const float* __restrict__ scal)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (scal) pTick = *scal;
    const u64 now = d.sc->clock, pass = d.sc->pass_index;
    const Philox4 r = philox4x32_10((u32)pass, (u32)(pass >> 32), i, STREAM_INJECT, kp.seed_lo, kp.seed_hi);
    if (u01_24(r.x) < pTick * v[i]) {
        d.view[i] = now;
        if (i >= kp.neuron_lo && i < kp.neuron_hi) d.live[i] = now;
    }
}

// Teacher forcing (brain-engine.cpp:126-133).
__global__ void k_teacher(const __grid_constant__ KParams kp, const DevPtrs d, const float* __restrict__ expected, u32 n,
                          float rate, u64 gap, const float* __restrict__ scal)
{
    const u32 o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    if (scal) rate = *scal;
    const u64 now = d.sc->clock, pass = d.sc->pass_index;
    const Philox4 r = philox4x32_10((u32)pass, (u32)(pass >> 32), o, STREAM_TEACHER, kp.seed_lo, kp.seed_hi);
    const float p = expected[o] * rate;
    const u64 id = (u64)kp.n_input + o;
    if (u01_24(r.x) < p && (now - d.view[id] > gap)) {
        d.view[id] = now;
        if (id >= kp.neuron_lo && id < kp.neuron_hi) d.live[id] = now;
    }
}

// abnn_engine_step prologue: inject_inputs + teacher forcing in one launch over the n_in + n_out head neurons
// (frame = [in | expected | pTick | rate] in device memory), and — when the exchange already delivered the
// gate words of this pass — the refresh of the head's words, which these two steps may just have changed.
__global__ void k_step_prologue(const __grid_constant__ KParams kp, const DevPtrs d, const float* __restrict__ frame,
                                u32 n_in, u32 n_out, u64 gap, u32 refresh_slack, u32 refresh_fire)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in + n_out) return;
    const u64 now = d.sc->clock, pass = d.sc->pass_index;
    const float pTick = frame[n_in + n_out], rate = frame[n_in + n_out + 1];
    u64 lp = d.view[i];
    bool spike;
    if (i < n_in) {                                                        // brain.cpp:73-83
        const Philox4 r = philox4x32_10((u32)pass, (u32)(pass >> 32), i, STREAM_INJECT, kp.seed_lo, kp.seed_hi);
        spike = u01_24(r.x) < pTick * frame[i];
    } else {                                                               // brain-engine.cpp:126-133
        const u32 o = i - n_in;
        const Philox4 r = philox4x32_10((u32)pass, (u32)(pass >> 32), o, STREAM_TEACHER, kp.seed_lo, kp.seed_hi);
        spike = u01_24(r.x) < frame[i] * rate && (now - lp > gap);
    }
    if (spike) {
        lp = now;
        d.view[i] = now;
        if (i >= kp.neuron_lo && i < kp.neuron_hi) d.live[i] = now;
    }
    if (refresh_slack) d.slack[i] = slack_word(now, lp, kp.window_pre);
    // k_traverse_line32's fire word of an owned head neuron (traversal.cu:k_fold_prepare32 wrote it before this step's spikes)
    if (refresh_fire && i >= kp.neuron_lo && i < kp.neuron_hi) d.fire32[i] = fire_word(now, d.live[i]);
}

// Brain::read_outputs (brain.cpp:145-157), window = ticks of the last pass.
__device__ __forceinline__ bool output_spiked(const KParams& kp, const DevPtrs& d, u32 o)
{
    const u64 now = d.sc->clock, span = d.sc->last_pass_ticks;
    const u64 start = now > span ? now - span : 0;
    const u64 ts = d.view[(u64)kp.n_input + o];
    return ts != 0 && ts >= start && ts < now;
}
__global__ void k_read_outputs(const __grid_constant__ KParams kp, const DevPtrs d, unsigned char* spikes, u32 n_out)
{
    const u32 o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o < n_out) spikes[o] = output_spiked(kp, d, o) ? 1 : 0;
}

// One CTA: rate EMA (brain-engine.cpp:145-151) -> RateFilter::process (rate-filter.h:22-59: IIR with a
// double-precision alpha, FIR = mean over the last <= fir_size IIR frames summed oldest-first in
// float) -> peak normalise (brain-engine.cpp:156-164) -> every reward_window calls the MSE loss in
// double, summed in index order, and reward = float(lastLoss - loss) (brain-engine.cpp:171-186).
__global__ void __launch_bounds__(256) k_readout(const __grid_constant__ KParams kp, const DevPtrs d, const ReadoutParams rp,
                                                 const ReadoutState rs, const float* __restrict__ expected)
{
    __shared__ float s_max[8];
    DevScalars* sc = d.sc;
    const u32 n = rp.n_output;
    const u32 count0 = sc->fir_count, head0 = sc->fir_head, init0 = sc->iir_init;
    const u32 count1 = rp.use_fir ? (count0 < rp.fir_size ? count0 + 1 : count0) : 0;
    const u32 head1 = rp.use_fir ? (count0 < rp.fir_size ? head0 : (head0 + 1) % rp.fir_size) : 0;
    const u32 slot = rp.use_fir ? (count0 < rp.fir_size ? (head0 + count0) % rp.fir_size : head0) : 0;
    float local_max = 0.f;
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
        const bool sp = output_spiked(kp, d, i);
        rs.spikes[i] = sp ? 1 : 0;
        const float r = (1.0f - rp.rate_alpha) * rs.rate[i] + rp.rate_alpha * (sp ? 1.f : 0.f);
        rs.rate[i] = r;
        float q = init0 ? rs.iir[i] : r;                                   // rate-filter.h:24-26
        q += (float)(rp.a * (double)(r - q));                              // rate-filter.h:32-34
        rs.iir[i] = q;
        float avg = q;
        if (rp.use_fir) {
            rs.fir[(size_t)slot * n + i] = q;                              // rate-filter.h:38-41
            avg = 0.f;
            for (u32 t = 0; t < count1; ++t) avg += rs.fir[(size_t)((head1 + t) % rp.fir_size) * n + i];
            avg *= 1.0f / (float)count1;                                   // rate-filter.h:50-53
        }
        rs.smooth[i] = avg;
        local_max = fmaxf(local_max, avg);
    }
    for (int o = 16; o; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = local_max;
    __syncthreads();
    float m = sc->max_observed;                                            // every thread reads before thread 0 writes
    for (u32 w = 0; w < (blockDim.x >> 5); ++w) m = fmaxf(m, s_max[w]);
    m *= rp.peak_decay;                                                    // brain-engine.cpp:160
    __syncthreads();
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) rs.smooth[i] = fminf(rs.smooth[i] / m, 1.0f);
    __syncthreads();
    if (threadIdx.x == 0) {
        sc->max_observed = m;
        sc->iir_init = 1; sc->fir_count = count1; sc->fir_head = head1;
        if (expected) {
            const u32 wp = sc->win_pos + 1;
            if (wp == rp.reward_window) {
                double loss = 0.0;
                for (u32 i = 0; i < n; ++i) { const double err = (double)(rs.smooth[i] - expected[i]); loss += err * err; }
                loss /= (double)n;
                sc->reward = (float)(sc->last_loss - loss);
                sc->last_loss = loss;
                sc->win_pos = 0;
                sc->windows_done += 1;
            } else sc->win_pos = wp;
        }
    }
}

cudaError_t launch_inject(const KParams& kp, const DevPtrs& d, const float* v, u32 n, float pTick, const float* scal, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    k_inject<<<(n + 255) / 256, 256, 0, st>>>(kp, d, v, n, pTick, scal);
    return cudaGetLastError();
}
cudaError_t launch_teacher(const KParams& kp, const DevPtrs& d, const float* expected, u32 n, float rate, u64 gap, const float* scal,
                           cudaStream_t st)
{
    if (!n) return cudaSuccess;
    k_teacher<<<(n + 255) / 256, 256, 0, st>>>(kp, d, expected, n, rate, gap, scal);
    return cudaGetLastError();
}
cudaError_t launch_step_prologue(const KParams& kp, const DevPtrs& d, const float* frame, u32 n_in, u32 n_out, u64 gap,
                                 bool refresh_slack, bool refresh_fire, cudaStream_t st)
{
    if (!(n_in + n_out)) return cudaSuccess;
    k_step_prologue<<<(n_in + n_out + 255) / 256, 256, 0, st>>>(kp, d, frame, n_in, n_out, gap, refresh_slack ? 1u : 0u,
                                                                refresh_fire ? 1u : 0u);
    return cudaGetLastError();
}
cudaError_t launch_read_outputs(const KParams& kp, const DevPtrs& d, unsigned char* spikes, u32 n_out, cudaStream_t st)
{
    if (!n_out) return cudaSuccess;
    k_read_outputs<<<(n_out + 255) / 256, 256, 0, st>>>(kp, d, spikes, n_out);
    return cudaGetLastError();
}
cudaError_t launch_readout(const KParams& kp, const DevPtrs& d, const ReadoutParams& rp, const ReadoutState& rs,
                           const float* expected, cudaStream_t st)
{
    k_readout<<<1, 256, 0, st>>>(kp, d, rp, rs, expected);
    return cudaGetLastError();
}

}  // namespace abnn
