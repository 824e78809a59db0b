// N3: Monte-Carlo SIR label generator -- the process of the reference's sir_torch (ode_nn.py:30-88).
//
// Reference: `sims` simulations ONE AFTER THE OTHER, each of T-1 steps of ~10 tiny tensor ops with torch.rand on the
// CPU (10^4 x 19 iterations for one label file). Per step, with idx_I the nodes infected at the START of the step:
//   every edge (u in idx_I, v susceptible) infects v with probability beta (independent coins, :59-66),
//   every u in idx_I recovers with probability gamma (:68-71),
//   then I[new_infected] = 1, I[new_recovered] = 0, S[new_infected] = 0, R[new_recovered] = 1 (:73-76),
// and the state is added to the per-time node counters (:78-80).
//
// Here: one CTA per simulation (grid-stride over the simulations of the call), state bit-packed in shared memory (S,
// I, newly-infected masks: 3 bits per node), counter-based Philox4x32-10 random numbers keyed by (seed; simulation,
// step, edge-or-node) so that the result does not depend on the grid or on scheduling, and EVENT histograms instead of
// per-step counter updates: a node's infection and recovery times are recorded once each (<= 2 atomics per node and
// simulation instead of 3 (T-1) additions), and the per-time counts follow by prefix sums in mc_finalize_kernel:
//   I_t = #infected by t - #recovered by t,  S_t = sims - #infected by t,  R_t = #recovered by t.
// Quirk kept: the reference ASSIGNS the t = 0 rows (I_per_sim[:,0,:] = I, :54-55) instead of accumulating them, so
// counts[., 0, .] are the 0/1 indicators of the initial state (the loss never reads t = 0).
#include <algorithm>

#include "gnode_common.cuh"

namespace gnode {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
// uniform in [0, 1) with 24 random bits: stream `kind` (0 = transmission coin of CSR entry idx, 1 = recovery coin of node idx)
__device__ __forceinline__ float mc_uniform(uint2 key, uint32_t sim, uint32_t step, uint32_t kind, uint32_t idx) {
    const uint4 r = philox4x32_10(make_uint4(idx >> 2, sim, step, kind), key);
    const uint32_t w = (idx & 3u) == 0 ? r.x : ((idx & 3u) == 1 ? r.y : ((idx & 3u) == 2 ? r.z : r.w));
    return (float)(w >> 8) * (1.0f / 16777216.0f);
}

struct McArgs {
    const int32_t* rowptr;
    const int32_t* colidx;
    const int32_t* seeds;
    int32_t n, n_seeds, T, sims, sim0;
    float beta, gamma;
    uint2 key;
    int32_t* inf_hist;   // [T][n] infection events at step t (t >= 1)
    int32_t* rec_hist;   // [T][n] recovery events at step t
};

__global__ void __launch_bounds__(256) mc_sir_kernel(const McArgs a) {
    extern __shared__ uint32_t masks[];
    const int words = (a.n + 31) >> 5;
    uint32_t* S = masks;
    uint32_t* I = masks + words;
    uint32_t* NI = masks + 2 * words;
    for (int sim = blockIdx.x; sim < a.sims; sim += gridDim.x) {
        const uint32_t sid = (uint32_t)(a.sim0 + sim);
        for (int w = threadIdx.x; w < words; w += blockDim.x) {
            const int rem = a.n - 32 * w;
            S[w] = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
            I[w] = 0u; NI[w] = 0u;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < a.n_seeds; i += blockDim.x) {
            const int s = a.seeds[i];
            if (s >= 0 && s < a.n) { atomicOr(&I[s >> 5], 1u << (s & 31)); atomicAnd(&S[s >> 5], ~(1u << (s & 31))); }
        }
        __syncthreads();
        for (int t = 1; t < a.T; ++t) {
            int any = 0;
            for (int w = threadIdx.x; w < words; w += blockDim.x) {
                uint32_t iw = I[w], rec = 0u;
                any |= iw != 0u;
                while (iw) {
                    const int bit = __ffs(iw) - 1;
                    iw &= iw - 1;
                    const int u = 32 * w + bit;
                    const int e1 = a.rowptr[u + 1];
                    for (int e = a.rowptr[u]; e < e1; ++e) {
                        const int v = a.colidx[e];
                        if ((S[v >> 5] >> (v & 31)) & 1u)
                            if (mc_uniform(a.key, sid, (uint32_t)t, 0u, (uint32_t)e) < a.beta) atomicOr(&NI[v >> 5], 1u << (v & 31));
                    }
                    if (mc_uniform(a.key, sid, (uint32_t)t, 1u, (uint32_t)u) < a.gamma) {
                        rec |= 1u << bit;
                        atomicAdd(&a.rec_hist[(size_t)t * a.n + u], 1);
                    }
                }
                if (rec) I[w] &= ~rec;          // only this thread touches word w of I in this phase
            }
            if (!__syncthreads_or(any)) break;  // the epidemic has died out: no further events
            for (int w = threadIdx.x; w < words; w += blockDim.x) {
                uint32_t nw = NI[w];
                if (nw) {
                    I[w] |= nw; S[w] &= ~nw; NI[w] = 0u;
                    while (nw) {
                        const int bit = __ffs(nw) - 1;
                        nw &= nw - 1;
                        atomicAdd(&a.inf_hist[(size_t)t * a.n + 32 * w + bit], 1);
                    }
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

// counts [3][T][n] (S, I, R) from the event histograms of all `sims_total` simulations
__global__ void __launch_bounds__(256) mc_finalize_kernel(const int32_t* __restrict__ inf_hist, const int32_t* __restrict__ rec_hist,
                                                          const int32_t* __restrict__ seeds, int n_seeds, int n, int T, int sims_total,
                                                          double* __restrict__ counts) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    bool seed = false;
    for (int i = 0; i < n_seeds; ++i) seed |= seeds[i] == v;
    const size_t plane = (size_t)T * n;
    counts[v] = seed ? 0.0 : 1.0;                      // t = 0: assigned, not accumulated (ode_nn.py:54-55)
    counts[plane + v] = seed ? 1.0 : 0.0;
    counts[2 * plane + v] = 0.0;
    long long inf = seed ? sims_total : 0, rec = 0;
    for (int t = 1; t < T; ++t) {
        inf += inf_hist[(size_t)t * n + v];
        rec += rec_hist[(size_t)t * n + v];
        counts[(size_t)t * n + v] = (double)(sims_total - inf);
        counts[plane + (size_t)t * n + v] = (double)(inf - rec);
        counts[2 * plane + (size_t)t * n + v] = (double)rec;
    }
}

}  // namespace gnode

using namespace gnode;

extern "C" size_t gnode_mc_sir_workspace_bytes(gnode_graph_t g, int32_t T) {
    if (!g || T < 1) return 0;
    return 2 * (size_t)T * (size_t)g->n * sizeof(int32_t);
}

extern "C" int gnode_mc_sir(gnode_graph_t g, const int32_t* seeds, int32_t n_seeds, float beta, float gamma, int32_t sims,
                            int32_t T, uint64_t rng_seed, double* counts, void* workspace, size_t workspace_bytes,
                            void* stream_) {
    if (!g || !seeds || n_seeds < 0 || sims < 1 || T < 1 || !counts || !workspace ||
        workspace_bytes < gnode_mc_sir_workspace_bytes(g, T)) {
        set_error("gnode_mc_sir: bad arguments (sims=%d T=%d n_seeds=%d)", sims, T, n_seeds);
        return GNODE_ERR_ARG;
    }
    int dev = -1;
    GN_CUDA(cudaGetDevice(&dev));
    if (dev != g->device) { set_error("gnode_mc_sir: the graph lives on device %d, current device is %d", g->device, dev); return GNODE_ERR_ARG; }
    const int words = (g->n + 31) / 32;
    const size_t smem = 3 * (size_t)words * sizeof(uint32_t);
    if (smem > 200 * 1024) {
        set_error("gnode_mc_sir: %d nodes need %zu bytes of bit-packed state per simulation (limit 200 KB of shared memory)", g->n, smem);
        return GNODE_ERR_UNSUPPORTED;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    if (smem > 48 * 1024) GN_CUDA(cudaFuncSetAttribute(mc_sir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    McArgs a;
    a.rowptr = g->d_rowptr; a.colidx = g->d_colidx; a.seeds = seeds;
    a.n = g->n; a.n_seeds = n_seeds; a.T = T; a.sims = sims; a.sim0 = 0;
    a.beta = beta; a.gamma = gamma;
    a.key = make_uint2((uint32_t)rng_seed, (uint32_t)(rng_seed >> 32));
    a.inf_hist = (int32_t*)workspace;
    a.rec_hist = a.inf_hist + (size_t)T * g->n;
    GN_CUDA(cudaMemsetAsync(workspace, 0, gnode_mc_sir_workspace_bytes(g, T), stream));
    int sm_count = 0;
    GN_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    const int threads = words <= 2 ? 32 : (words <= 64 ? 64 : 256);          // one thread per word of the masks, up to a CTA
    const int per_sm = std::max(1, std::min(32, (int)((200 * 1024) / std::max<size_t>(smem, 1024))));
    const int grid = std::min(sims, sm_count * std::min(per_sm, 2048 / threads));
    mc_sir_kernel<<<grid, threads, smem, stream>>>(a);
    GN_LAUNCH_CHECK();
    mc_finalize_kernel<<<(g->n + 255) / 256, 256, 0, stream>>>(a.inf_hist, a.rec_hist, seeds, n_seeds, g->n, T, sims, counts);
    GN_LAUNCH_CHECK();
    return GNODE_OK;
}
