// Forward rollout: encoder, fused Euler-step kernel, decoder (a1-a9).
//
// One CTA owns a tile of 128 consecutive rows of the batch ((trial, node) pairs) and
// runs, per Euler step k (reference lines in brackets, ode_nn_ngraph_sim.py):
//   1. S_k tile  -> shared memory (UMMA K-major SWIZZLE_128B layout)
//   2. S' = sigmoid(S_k W^T + b)                        [:62-63]   -> shared memory
//   3. row-per-half-warp: AI = sum_nbrs I'_k            [:73]      (I'_k from HBM/L2)
//        dS,dI,dR ; y_{k+1} = y_k + dt f                [:75-77,96; torchdiffeq Euler]
//        -> HBM (coalesced); I_{k+1} stays in shared memory; decoder + softmax of
//        y_{k+1} -> probs[k+1]                          [:170-188]
//   4. I'_{k+1} = sigmoid(I_{k+1} W^T + b) -> HBM       (next step's aggregation operand)
// so one launch per step reads S,I,R,I' and writes S,I,R,I',probs exactly once; the
// grid-wide dependency (all of I'_k before any aggregation) is the launch boundary.
// The reference's R' = sigmoid(linear(R)) is never used (:66 vs :75-77) and is skipped.
#include <algorithm>
#include <cstdlib>

#include "gnode_common.cuh"
#include "gnode_tile.cuh"
#include "gnode_umma.cuh"

#ifndef GNODE_DEFAULT_VARIANT
#define GNODE_DEFAULT_VARIANT 0
#endif

namespace gnode {

enum { MODE_STEP = 0, MODE_ENCODE = 1, MODE_IP = 2, MODE_RHS = 3 };

struct StepArgs {
    GnBatchView bv;
    const float* y_in;    // [3][M][H]
    float* y_out;         // [3][M][H]  STEP: y_{k+1}; ENCODE: y_0; RHS: f(y)
    const float* ip_in;   // [M][H]     I'_k
    float* ip_out;        // [M][H]     I'_{k+1}
    float* beta;          // [M]
    float* gamma;         // [M]
    const float* x;       // ENCODE: [M][ldx]
    int64_t ldx;
    float* probs;         // [M][3] slice of the produced state, or null
    float dt;
    int* counter;         // dynamic tile scheduler (one zeroed int per launch) or null = static striding
    gnode_params_t p;
};

// shared-memory carve-up (bytes from a 1024-B aligned base; operand tiles need 1024-B alignment)
constexpr int SM_X = 0;                        // 32 KB  operand tile (S_k, then I_{k+1})
constexpr int SM_SP = 32768;                   // 32 KB  Xlo scratch / S' tile / I'_{k+1} staging
constexpr int SM_W = 65536;                    // 16 KB  FFMA: W [h][k] row-major; tensor path: Whi operand
constexpr int SM_WLO = SM_W + H * H * 4;       // 16 KB  tensor path: Wlo operand
constexpr int SM_B = SM_WLO + H * H * 4;       // bias [64]
constexpr int SM_W3 = SM_B + H * 4;            // linear3.weight [4][64]
constexpr int SM_W1 = SM_W3 + 4 * H * 4;       // linearS1.weight [64]
constexpr int SM_B1 = SM_W1 + H * 4;           // linearS1.bias [64]
constexpr int SM_SMALL = SM_B1 + H * 4;        // b3[4], w2[4], b2[1]
constexpr int SM_MBAR = SM_SMALL + 64;         // mbarrier (8 B) + TMEM base slot (4 B)
constexpr int SM_TOTAL = SM_MBAR + 16 + 1024;  // + slack for the manual 1024-B alignment

// kernel variants: bit 0 = tcgen05 3xTF32 transform (else FFMA), bit 1 = MUFU sigmoid (else expf + IEEE div)
constexpr int VAR_TC = 1, VAR_FASTSIG = 2;

// decoder + softmax of one row held 4 channels per lane by a half-warp
// (linear3 -> ReLU -> linearS2 -> softmax over {S,I,R}; ode_nn_ngraph_sim.py:172-187)
__device__ __forceinline__ void decode_row(float4 s, float4 i, float4 r, const float* W3s, const float* small,
                                           int l, bool valid, float* probs_row) {
    float v[12];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const float4 w = *reinterpret_cast<const float4*>(W3s + m * H + 4 * l);
        v[m] = dot4(s, w); v[4 + m] = dot4(i, w); v[8 + m] = dot4(r, w);
    }
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1)
#pragma unroll
        for (int m = 0; m < 12; ++m) v[m] += __shfl_xor_sync(0xffffffffu, v[m], off);
    if (l == 0 && valid) {
        const float* b3 = small; const float* w2 = small + 4; const float b2 = small[8];
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float acc = b2;
#pragma unroll
            for (int m = 0; m < 4; ++m) acc = fmaf(w2[m], fmaxf(v[4 * c + m] + b3[m], 0.f), acc);
            o[c] = acc;
        }
        const float mx = fmaxf(o[0], fmaxf(o[1], o[2]));
        const float e0 = expf(o[0] - mx), e1 = expf(o[1] - mx), e2 = expf(o[2] - mx);
        const float inv = 1.0f / (e0 + e1 + e2);
        probs_row[0] = e0 * inv; probs_row[1] = e1 * inv; probs_row[2] = e2 * inv;
    }
}

template <int MODE, int VAR>
__global__ void __launch_bounds__(NTHREADS, 2) step_kernel(const StepArgs a) {
    constexpr bool TC = (VAR & VAR_TC) != 0;
    constexpr bool FAST = (VAR & VAR_FASTSIG) != 0;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* Xs = smem + SM_X;
    unsigned char* SPs = smem + SM_SP;
    float* Ws = reinterpret_cast<float*>(smem + SM_W);
    float* bs = reinterpret_cast<float*>(smem + SM_B);
    float* W3s = reinterpret_cast<float*>(smem + SM_W3);
    float* w1s = reinterpret_cast<float*>(smem + SM_W1);
    float* b1s = reinterpret_cast<float*>(smem + SM_B1);
    float* small = reinterpret_cast<float*>(smem + SM_SMALL);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + SM_MBAR);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + SM_MBAR + 8);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int l = tid & 15;           // 16-byte chunk of the row owned in the row-per-half-warp phases
    const int hw = tid >> 4;          // half-warp id 0..31
    const int M = a.bv.M;
    const size_t plane = (size_t)M * H;

    // parameters -> shared memory (once per CTA; the CTA is persistent over its tiles)
    umma::Ctx cx;
    if (TC) {
        umma::prepare_weights(a.p.lin_w, smem + SM_W, smem + SM_WLO, tid, NTHREADS);
        if (tid < 32) umma::tmem_alloc(tslot, umma::TMEM_COLS);          // warp 0 owns alloc / dealloc
        if (tid == 0) umma::mbar_init(mbar, 1);
        umma::fence_before_sync();
    } else {
        for (int i = tid; i < H * H / 4; i += NTHREADS)
            reinterpret_cast<float4*>(Ws)[i] = reinterpret_cast<const float4*>(a.p.lin_w)[i];
    }
    if (tid < H) bs[tid] = a.p.lin_b[tid];
    if (MODE == MODE_ENCODE && tid < H) { w1s[tid] = a.p.s1_w[tid]; b1s[tid] = a.p.s1_b[tid]; }
    if (MODE == MODE_ENCODE || MODE == MODE_STEP) {     // decoder weights (RHS / IP callers pass none)
        if (tid < 4 * H) W3s[tid] = a.p.l3_w[tid];
        if (tid < 4) { small[tid] = a.p.l3_b[tid]; small[4 + tid] = a.p.s2_w[tid]; }
        if (tid == 0) small[8] = a.p.s2_b[0];
    }
    __syncthreads();
    if (TC) {
        umma::fence_after_sync();
        cx.tmem = *tslot;
        cx.bar = mbar;
        cx.phase = 0;
        cx.whi = umma::smem_u32(smem + SM_W);
        cx.wlo = umma::smem_u32(smem + SM_WLO);
    }

    int* tile_slot = reinterpret_cast<int*>(smem + SM_MBAR + 12);
    for (int seq = blockIdx.x;; seq += gridDim.x) {
        if (a.counter != nullptr) {               // dynamic: next entry of the hub-first processing order
            if (tid == 0) *tile_slot = atomicAdd(a.counter, 1);
            __syncthreads();
            seq = *tile_slot;
        }
        if (seq >= a.bv.n_tiles) break;
        const int tile = a.bv.tile_order[seq];
        const int64_t tile0 = (int64_t)tile * TILE;

        if (MODE == MODE_STEP || MODE == MODE_RHS) {
            load_tile(Xs, a.y_in, tile0, M, tid);                 // S_k
            __syncthreads();
            if (TC) umma::gemm_sigmoid_tc<FAST>(cx, Xs, SPs, bs, tid);   // S'
            else gemm_sigmoid<FAST>(Xs, Ws, bs, SPs, tid);
            __syncthreads();
        } else if (MODE == MODE_IP) {
            load_tile(Xs, a.y_in + plane, tile0, M, tid);         // I
            __syncthreads();
        }

        if (MODE == MODE_STEP || MODE == MODE_RHS) {
            int inst = a.bv.tile_inst[tile];
#pragma unroll 1
            for (int it = 0; it < TILE / 32; ++it) {
                const int rr = hw + 32 * it;
                const int64_t g = tile0 + rr;
                const bool valid = g < M;
                int row0 = 0, e0 = 0, deg = 0;
                const int32_t* ci = nullptr;
                if (valid) {
                    while (inst + 1 < a.bv.n_inst && a.bv.inst[inst + 1].row0 <= g) ++inst;
                    const GnInstance I = a.bv.inst[inst];
                    row0 = I.row0; ci = I.colidx;
                    const int n = (int)(g - row0);
                    e0 = I.rowptr[n];
                    deg = I.rowptr[n + 1] - e0;
                }
                // ---- aggregation: sequential ascending-column sum (ode_nn_ngraph_sim.py:73)
                const float4 acc = gather_row(a.ip_in, ci, e0, deg, row0, l, lane);
                // ---- SIR derivative + Euler update (explicit _rn ops: no FMA contraction, the
                //      reference rounds after every ATen op)
                float4 sn = make_float4(0.f, 0.f, 0.f, 0.f), in_ = sn, rn = sn;
                if (valid) {
                    const size_t off = (size_t)g * H + 4 * l;
                    const float4 ipo = ldg4(a.ip_in + off);
                    const float4 s = lds4(Xs, sw_off(rr, l));
                    const float4 sp = lds4(SPs, sw_off(rr, l));
                    const float4 iv = ldg4_stream(a.y_in + plane + off);
                    const float4 rv = ldg4_stream(a.y_in + 2 * plane + off);
                    const float nbe = -a.beta[g], ga = a.gamma[g], dt = a.dt;
#define GN_COMP(c)                                                                  \
    {                                                                               \
        const float dS = __fmul_rn(nbe, __fmul_rn(acc.c, sp.c));                    \
        const float dR = __fmul_rn(ga, ipo.c);                                      \
        const float dI = __fsub_rn(-dS, dR);                                        \
        if (MODE == MODE_RHS) { sn.c = dS; in_.c = dI; rn.c = dR; }                 \
        else {                                                                      \
            sn.c = __fadd_rn(s.c, __fmul_rn(dt, dS));                               \
            in_.c = __fadd_rn(iv.c, __fmul_rn(dt, dI));                             \
            rn.c = __fadd_rn(rv.c, __fmul_rn(dt, dR));                              \
        }                                                                           \
    }
                    GN_COMP(x) GN_COMP(y) GN_COMP(z) GN_COMP(w)
#undef GN_COMP
                    stg4_stream(a.y_out + off, sn);
                    stg4_stream(a.y_out + plane + off, in_);
                    stg4_stream(a.y_out + 2 * plane + off, rn);
                    if (MODE == MODE_STEP) sts4(Xs, sw_off(rr, l), in_);      // operand of the next GEMM
                }
                if (MODE == MODE_STEP && a.probs != nullptr)
                    decode_row(sn, in_, rn, W3s, small, l, valid, a.probs + (size_t)(valid ? g : 0) * 3);
            }
            __syncthreads();
        } else if (MODE == MODE_ENCODE) {
            // encoder: C0 = relu(c * w1 + b1) for c in {S0, I0, R0}  (ode_nn_ngraph_sim.py:151-156)
#pragma unroll 1
            for (int it = 0; it < TILE / 32; ++it) {
                const int rr = hw + 32 * it;
                const int64_t g = tile0 + rr;
                const bool valid = g < M;
                float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), i0 = s0, r0 = s0;
                if (valid) {
                    const float* xr = a.x + (size_t)g * a.ldx;
                    const float cs = xr[0], ci_ = xr[1], cr = xr[2];
                    if (l == 0) { a.beta[g] = xr[3]; a.gamma[g] = xr[4]; }
                    const float4 w = *reinterpret_cast<const float4*>(w1s + 4 * l);
                    const float4 b = *reinterpret_cast<const float4*>(b1s + 4 * l);
#define GN_ENC(c)                                                             \
    s0.c = fmaxf(__fadd_rn(__fmul_rn(cs, w.c), b.c), 0.f);                    \
    i0.c = fmaxf(__fadd_rn(__fmul_rn(ci_, w.c), b.c), 0.f);                   \
    r0.c = fmaxf(__fadd_rn(__fmul_rn(cr, w.c), b.c), 0.f);
                    GN_ENC(x) GN_ENC(y) GN_ENC(z) GN_ENC(w)
#undef GN_ENC
                    const size_t off = (size_t)g * H + 4 * l;
                    stg4_stream(a.y_out + off, s0);
                    stg4_stream(a.y_out + plane + off, i0);
                    stg4_stream(a.y_out + 2 * plane + off, r0);
                }
                sts4(Xs, sw_off(rr, l), i0);
                if (a.probs != nullptr)
                    decode_row(s0, i0, r0, W3s, small, l, valid, a.probs + (size_t)(valid ? g : 0) * 3);
            }
            __syncthreads();
        }

        if (MODE != MODE_RHS) {
            if (TC) umma::gemm_sigmoid_tc<FAST>(cx, Xs, SPs, bs, tid);   // I'_{k+1}
            else gemm_sigmoid<FAST>(Xs, Ws, bs, SPs, tid);
            __syncthreads();
            store_tile(a.ip_out, SPs, tile0, M, tid);
            __syncthreads();
        }
    }
    if (TC) {
        umma::fence_before_sync();
        __syncthreads();
        if (tid < 32) umma::tmem_dealloc(cx.tmem, umma::TMEM_COLS);
    }
}

// 0 = FFMA + accurate sigmoid ... 3 = tcgen05 + MUFU sigmoid; chosen by gnode_set_variant() / GNODE_VARIANT
static int g_variant = -1;

static int current_variant() {
    if (g_variant < 0) {
        const char* e = getenv("GNODE_VARIANT");
        g_variant = e ? (atoi(e) & 3) : GNODE_DEFAULT_VARIANT;
    }
    return g_variant;
}

template <int MODE, int VAR>
static int launch_step_v(const gnode_batch* b, const StepArgs& a, cudaStream_t stream) {
    static bool configured[64] = {false};
    if (!configured[b->device & 63]) {
        GN_CUDA(cudaFuncSetAttribute(step_kernel<MODE, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        configured[b->device & 63] = true;
    }
    const int grid = std::min(b->n_tiles, 2 * b->sm_count);
    step_kernel<MODE, VAR><<<grid, NTHREADS, SM_TOTAL, stream>>>(a);
    GN_LAUNCH_CHECK();
    return GNODE_OK;
}

template <int MODE>
static int launch_step(const gnode_batch* b, const StepArgs& a, cudaStream_t stream) {
    switch (current_variant()) {
        case 0: return launch_step_v<MODE, 0>(b, a, stream);
        case 1: return launch_step_v<MODE, 1>(b, a, stream);
        case 2: return launch_step_v<MODE, 2>(b, a, stream);
        default: return launch_step_v<MODE, 3>(b, a, stream);
    }
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace gnode

using namespace gnode;

extern "C" int gnode_set_variant(int variant) {
    if (variant < 0 || variant > 3) { set_error("gnode_set_variant: variant must be 0..3"); return GNODE_ERR_ARG; }
    g_variant = variant;
    return GNODE_OK;
}
extern "C" int gnode_get_variant(void) { return current_variant(); }

extern "C" size_t gnode_rollout_workspace_bytes(gnode_batch_t b, int with_traj) {
    if (!b) return 0;
    const size_t M = (size_t)b->M;
    size_t bytes = 2 * align_up(M * sizeof(float), 256);          // beta, gamma
    bytes += 4096;                                                // tile-scheduler counters (one int per launch)
    bytes += 2 * align_up(M * H * sizeof(float), 256);            // I' ping-pong
    if (!with_traj) bytes += 2 * align_up(3 * M * H * sizeof(float), 256);  // state ping-pong
    return bytes;
}

extern "C" int gnode_rollout_forward(gnode_batch_t b, const float* x, int64_t ldx, const gnode_params_t* p,
                                     int32_t T, const float* dt_host, float* traj, float* probs,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
    if (!b || !x || !p || !probs || !workspace || T < 1 || ldx < 5 || (T > 1 && !dt_host)) {
        set_error("gnode_rollout_forward: bad arguments (T=%d ldx=%lld)", T, (long long)ldx);
        return GNODE_ERR_ARG;
    }
    if (workspace_bytes < gnode_rollout_workspace_bytes(b, traj != nullptr)) {
        set_error("gnode_rollout_forward: workspace too small (%zu < %zu)", workspace_bytes,
                  gnode_rollout_workspace_bytes(b, traj != nullptr));
        return GNODE_ERR_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t M = (size_t)b->M;
    unsigned char* ws = (unsigned char*)workspace;
    float* beta = (float*)ws;  ws += align_up(M * sizeof(float), 256);
    float* gamma = (float*)ws; ws += align_up(M * sizeof(float), 256);
    int* counters = (int*)ws;  ws += 4096;
    GN_CUDA(cudaMemsetAsync(counters, 0, 4096, stream));
    float* ip[2];
    ip[0] = (float*)ws; ws += align_up(M * H * sizeof(float), 256);
    ip[1] = (float*)ws; ws += align_up(M * H * sizeof(float), 256);
    float* st[2] = {nullptr, nullptr};
    if (!traj) {
        st[0] = (float*)ws; ws += align_up(3 * M * H * sizeof(float), 256);
        st[1] = (float*)ws;
    }
    auto state = [&](int k) -> float* { return traj ? traj + (size_t)k * 3 * M * H : st[k & 1]; };

    StepArgs a;
    a.bv = gn_view(b);
    a.p = *p;
    a.beta = beta; a.gamma = gamma;
    a.x = x; a.ldx = ldx;
    a.y_in = nullptr; a.ip_in = nullptr;
    a.y_out = state(0); a.ip_out = ip[0];
    a.probs = probs; a.dt = 0.f;
    a.counter = counters;
    int rc = launch_step<MODE_ENCODE>(b, a, stream);
    if (rc) return rc;
    for (int k = 0; k + 1 < T; ++k) {
        a.y_in = state(k); a.y_out = state(k + 1);
        a.ip_in = ip[k & 1]; a.ip_out = ip[(k + 1) & 1];
        a.probs = probs + (size_t)(k + 1) * M * 3;
        a.dt = dt_host[k];
        a.counter = (k + 1 < 1024) ? counters + (k + 1) : nullptr;
        rc = launch_step<MODE_STEP>(b, a, stream);
        if (rc) return rc;
    }
    return GNODE_OK;
}

extern "C" int gnode_odefunc_eval(gnode_batch_t b, const float* y, const float* beta, const float* gamma,
                                  const gnode_params_t* p, float* dy, float* scratch, void* stream_) {
    if (!b || !y || !beta || !gamma || !p || !dy || !scratch) {
        set_error("gnode_odefunc_eval: null argument");
        return GNODE_ERR_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    StepArgs a;
    a.bv = gn_view(b);
    a.p = *p;
    a.beta = const_cast<float*>(beta); a.gamma = const_cast<float*>(gamma);
    a.x = nullptr; a.ldx = 0; a.probs = nullptr; a.dt = 0.f; a.counter = nullptr;
    a.y_in = y; a.y_out = nullptr; a.ip_in = nullptr; a.ip_out = scratch;
    int rc = launch_step<MODE_IP>(b, a, stream);       // I' of every row first (grid-wide dependency)
    if (rc) return rc;
    a.ip_in = scratch; a.ip_out = nullptr; a.y_out = dy;
    return launch_step<MODE_RHS>(b, a, stream);
}
