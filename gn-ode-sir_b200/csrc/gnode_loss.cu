// N1: L1 loss of the sub-sampled prediction against the Monte-Carlo labels and its cotangent, one pass.
//
// The reference's train()/test() (ode_nn_ngraph_sim.py:230-234, ode_nn_ngraphs.py:216-220) copy the rows
// x_rk[int(i/deltaT)] of S, I and R one by one into CPU tensors (get_sir_t_nodes_torch, ode_nn.py:249-261: 3 x maxTime
// device -> host copies per mini-batch), cat / transpose them, move the result back to the device and apply
// nn.L1Loss()(pred[:,1:,:], y[:,1:,:]) with float64 labels (the prediction is promoted). Here the rollout already emits
// only the consumed grid points (gnode_rollout_forward_sel), and this kernel reads them once:
//   loss = mean_{m, t >= skip, c} |probs[t,m,c] - labels[m,t,c]|,  grad[t,m,c] = sign(.) * scale / count (0 for t < skip).
// Layouts: probs / grad time-major [n_out,M,3] fp32 (as the rollout writes / the reverse sweep reads them), labels
// node-major [M,n_out,3] fp64 (the reference's y.view(-1, maxTime, 3)). A block stages the labels of a tile of nodes in
// shared memory (coalesced in the labels' order) and then walks the tile time-major (coalesced in the probabilities'
// order). float64 accumulation in a fixed order: per-thread, block tree, then one block over the per-block slots.
#include <algorithm>

#include "gnode_common.cuh"

namespace gnode {

constexpr int L1_BLOCKS = 1024, L1_THREADS = 256;
constexpr int L1_SMEM_DOUBLES = 5760;             // 45 KB of staged labels per block

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < L1_THREADS / 32; ++i) s += red[i];
    __syncthreads();
    return s;                                      // valid in thread 0
}

__global__ void __launch_bounds__(L1_THREADS) l1_partial_kernel(const float* __restrict__ probs, const double* __restrict__ labels,
                                                                int64_t M, int n_out, int skip, int tile_nodes, float g,
                                                                double* __restrict__ part, float* __restrict__ grad) {
    __shared__ double lab_s[L1_SMEM_DOUBLES];
    __shared__ double red[L1_THREADS / 32];
    const int64_t n_tiles = (M + tile_nodes - 1) / tile_nodes;
    const int row3 = n_out * 3;
    double acc = 0.0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * tile_nodes;
        const int nm = (int)min((int64_t)tile_nodes, M - m0);
        const double* lsrc = labels + (size_t)m0 * row3;
        for (int i = threadIdx.x; i < nm * row3; i += L1_THREADS) lab_s[i] = lsrc[i];
        __syncthreads();
        const int per_t = nm * 3;
        for (int i = threadIdx.x; i < n_out * per_t; i += L1_THREADS) {
            const int t = i / per_t, r = i - t * per_t;          // r = 3 * m_local + c
            const size_t pi = ((size_t)t * M + m0) * 3 + r;
            float go = 0.f;
            if (t >= skip) {
                const int ml = r / 3, c = r - 3 * ml;
                const double d = (double)probs[pi] - lab_s[(ml * n_out + t) * 3 + c];
                acc += fabs(d);
                go = d > 0.0 ? g : (d < 0.0 ? -g : 0.f);
            }
            if (grad) grad[pi] = go;
        }
        __syncthreads();
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(L1_THREADS) l1_final_kernel(const double* __restrict__ part, int n, double inv_count,
                                                              double* __restrict__ loss_out) {
    __shared__ double red[L1_THREADS / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += L1_THREADS) acc += part[i];
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) loss_out[0] = s * inv_count;
}

}  // namespace gnode

using namespace gnode;

extern "C" size_t gnode_l1_scratch_bytes(void) { return L1_BLOCKS * sizeof(double); }

extern "C" int gnode_l1_loss_grad(const float* probs, const double* labels, int64_t M, int32_t n_out, int32_t skip,
                                  float scale, double* loss_out, float* grad_probs, void* scratch, void* stream_) {
    if (!probs || !labels || !loss_out || !scratch || M < 1 || n_out < 1 || skip < 0 || skip >= n_out) {
        set_error("gnode_l1_loss_grad: bad arguments (M=%lld n_out=%d skip=%d)", (long long)M, n_out, skip);
        return GNODE_ERR_ARG;
    }
    if (3 * n_out > L1_SMEM_DOUBLES) {
        set_error("gnode_l1_loss_grad: n_out = %d exceeds %d", n_out, L1_SMEM_DOUBLES / 3);
        return GNODE_ERR_UNSUPPORTED;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int tile_nodes = std::min(64, L1_SMEM_DOUBLES / (3 * n_out));
    const int64_t n_tiles = (M + tile_nodes - 1) / tile_nodes;
    const int grid = (int)std::min<int64_t>(n_tiles, L1_BLOCKS);
    const double count = (double)M * (double)(n_out - skip) * 3.0;
    l1_partial_kernel<<<grid, L1_THREADS, 0, stream>>>(probs, labels, M, n_out, skip, tile_nodes, (float)((double)scale / count),
                                                       (double*)scratch, grad_probs);
    GN_LAUNCH_CHECK();
    l1_final_kernel<<<1, L1_THREADS, 0, stream>>>((const double*)scratch, grid, 1.0 / count, loss_out);
    GN_LAUNCH_CHECK();
    return GNODE_OK;
}
