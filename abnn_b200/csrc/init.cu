// abnn_b200/csrc/init.cu — device-side graph initialisation (README.md:134-135: Erdős–Rényi endpoints,
// weights ~ Beta(2,8)). Edge g is a pure function of (seed, g): four Philox calls give src, dst and
// nine uniforms; w is their 2nd smallest (the a-th order statistic of a+b-1 uniforms is Beta(a,b)),
// so the table needs no transcendental functions and is reproducible for a given (seed, world_size): rank r generates
// the edge ids [n*r/W, n*(r+1)/W) with dst uniform over ITS OWN neuron slice, i.e. the graph — and with it every result
// of a sharded run — depends on the world size (a W-rank run is compared with the W-shard oracle, never with the
// single-GPU run; include/abnn.h, "GPU-count invariance").
// The reference's own build_random_graph (brain-engine.cpp:31-53) depends on the host C++ library's
// mt19937 distributions and is therefore generated on the host (capi.cu) and uploaded.
#include "common.cuh"
#include "kernels.h"

namespace abnn {

__global__ void __launch_bounds__(256) k_init_er_beta(abnn_synapse* syn, u64 g0, u64 count, u32 seed_lo, u32 seed_hi,
                                                      u64 n_neuron, u64 dlo, u64 dspan)
{
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (u64)gridDim.x * blockDim.x) {
        const u64 g = g0 + t;
        u32 r[16];
#pragma unroll
        for (u32 c = 0; c < 4; ++c) {
            const Philox4 q = philox4x32_10((u32)g, (u32)(g >> 32), c, STREAM_INIT, seed_lo, seed_hi);
            r[4 * c] = q.x; r[4 * c + 1] = q.y; r[4 * c + 2] = q.z; r[4 * c + 3] = q.w;
        }
        const u32 src = (u32)mulhi64(((u64)r[0] << 32) | r[1], n_neuron);
        const u32 dst = (u32)(dlo + mulhi64(((u64)r[2] << 32) | r[3], dspan));
        float m1 = 2.f, m2 = 2.f;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float u = u01_24(r[4 + j]);
            if (u < m1) { m2 = m1; m1 = u; } else if (u < m2) { m2 = u; }
        }
        reinterpret_cast<uint4*>(syn)[t] = make_uint4(src, dst, __float_as_uint(m2), 0u);
    }
}

cudaError_t launch_init_er_beta(abnn_synapse* syn, u64 g0, u64 count, u64 seed, u64 n_neuron, u64 dlo, u64 dhi,
                                int sm_count, cudaStream_t st)
{
    if (!count) return cudaSuccess;
    u64 blocks = (count + 255) / 256;
    const u64 cap = (u64)sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_init_er_beta<<<(unsigned)blocks, 256, 0, st>>>(syn, g0, count, (u32)seed, (u32)(seed >> 32), n_neuron, dlo, dhi - dlo);
    return cudaGetLastError();
}

}  // namespace abnn
