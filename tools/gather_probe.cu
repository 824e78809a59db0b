// Probe: neighbour aggregation AI[r] = sum_{c in N(r)} X[c] over a BA graph replicated over B trials,
// (a) with per-lane LDG gathers (the round-1 step kernel's scheme) and (b) with a TMA tile::gather4 producer
// warp filling a shared-memory ring that half-warp-per-row consumers sum in ascending-column order.
// mode "skel" adds the step kernel's own-row streaming traffic (4 row reads + 4 row writes) to both, which
// gives the memory-system ceiling of the fused Euler step with all arithmetic removed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/gather_probe tools/gather_probe.cu
//   tools/gather_probe [trials=64] [nst=3] [reps=5]
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int H = 64, TILE = 128, SE = 64;          // SE = ring entries (rows of 256 B) per stage
constexpr int NCW = 16;                              // consumer warps
constexpr int RPN = 8;                               // rowptr-slice slots (>= max stages + 1)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 22)) { printf("mbar timeout blk %d thr %d bar %u par %u\n", blockIdx.x, threadIdx.x, addr, parity); __trap(); }
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t pol_evict_last() { uint64_t p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_evict_first() { uint64_t p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ float4 ldg4_hint(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ldg4_na(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg4_hint(float* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

struct Args {
    const int* rowptr;     // [N+1]
    const int* colidx;     // [nnz]
    int N, B, blocks_per_trial, n_tiles;
    const float* X;        // [M][H]  (gather operand, "I'")
    const float* Y;        // [3][M][H] own-row streams (skel)
    float* out;            // [M][H]
    float* Yout;           // [3][M][H] (skel)
    int* counter;
    int skel;
};

// ------------------------------------------------------------------ (a) LDG gather
template <int BATCH, int MINB, int NA = 0, int NSM = 0>
__global__ void __launch_bounds__(512, MINB) gather_ldg(const Args a) {
    __shared__ int tile_s;
    __shared__ float4 pad_s[NSM > 0 ? 512 : 1];
    extern __shared__ unsigned char dyn_smem[];
    if (a.skel == 77) dyn_smem[threadIdx.x] = 1;      // keep the dynamic allocation alive
    const int tid = threadIdx.x, lane = tid & 31, l = tid & 15, hw = tid >> 4;
    const uint64_t keep = pol_evict_last(), stream = pol_evict_first();
    const size_t plane = (size_t)a.N * a.B * H;
    for (;;) {
        if (tid == 0) tile_s = atomicAdd(a.counter, 1);
        __syncthreads();
        const int tile = tile_s;
        __syncthreads();
        if (tile >= a.n_tiles) break;
        const int b = tile / a.blocks_per_trial, blk = tile - b * a.blocks_per_trial;
        const int n0 = blk * TILE, nrows = min(TILE, a.N - n0);
        const int row0 = b * a.N;
        const float* lane_base = a.X + 4 * l;
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
            const int rr = hw + 32 * it;
            int e0 = 0, deg = 0;
            if (rr < nrows) { e0 = a.rowptr[n0 + rr]; deg = a.rowptr[n0 + rr + 1] - e0; }
            const size_t off = (size_t)(row0 + n0 + min(rr, nrows - 1)) * H + 4 * l;
            float4 o0, o1, o2, o3;
            if (a.skel) {
                if (NA) { o0 = ldg4_na(a.X + off, keep); o1 = ldg4_na(a.Y + off, stream); o2 = ldg4_na(a.Y + plane + off, stream); o3 = ldg4_na(a.Y + 2 * plane + off, stream); }
                else {
                o0 = ldg4_hint(a.X + off, keep);
                o1 = ldg4_hint(a.Y + off, stream); o2 = ldg4_hint(a.Y + plane + off, stream); o3 = ldg4_hint(a.Y + 2 * plane + off, stream);
                }
            }
            const int degm = max(deg, __shfl_xor_sync(0xffffffffu, deg, 16));
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = 0; j < degm; j += BATCH) {
                int c[BATCH];
#pragma unroll
                for (int k = 0; k < BATCH; ++k) c[k] = (j + k < deg) ? a.colidx[e0 + j + k] + row0 : -1;
                float4 v[BATCH];
#pragma unroll
                for (int k = 0; k < BATCH; ++k) v[k] = (c[k] >= 0) ? (NA ? ldg4_na(lane_base + (size_t)c[k] * H, keep) : ldg4_hint(lane_base + (size_t)c[k] * H, keep)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < BATCH; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
            }
            if (NSM > 0) {          // extra shared-memory wavefronts: NSM 128-bit accesses per row (LSU pipe pressure test)
                volatile float4* ps = pad_s;
#pragma unroll
                for (int i = 0; i < NSM; i += 2) {
                    ps[tid].x = acc.x + (float)i;
                    acc.y += ps[tid ^ 16].x * 1e-30f;
                }
                float4 tmp; tmp.x = acc.x; tmp.y = acc.y; tmp.z = acc.z; tmp.w = acc.w;
                for (int i = 0; i < NSM; i += 2) { *(float4*)&pad_s[tid] = tmp; __syncwarp(); tmp = *(float4*)&pad_s[tid ^ 1]; __syncwarp(); }
                acc.z += tmp.z * 1e-30f;
            }
            if (rr < nrows) {
                if (a.skel) {
                    stg4_hint(a.Yout + off, make_float4(o1.x + acc.x, o1.y + acc.y, o1.z, o1.w), stream);
                    stg4_hint(a.Yout + plane + off, make_float4(o2.x + o0.x, o2.y, o2.z + acc.z, o2.w), stream);
                    stg4_hint(a.Yout + 2 * plane + off, make_float4(o3.x, o3.y + o0.y, o3.z, o3.w + acc.w), stream);
                }
                stg4_hint(a.out + off, acc, stream);
            }
        }
    }
}

// ------------------------------------------------------------------ (b) TMA gather4 ring
// smem: ring[NST][SE][256 B] | rp_s[RPN][TILE+4] | tq[RPN] | full[NST], empty[NST]
template <int NST>
__global__ void __launch_bounds__((NCW + 1) * 32, 1) gather_ring(const Args a, const __grid_constant__ CUtensorMap tmX) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* ring = reinterpret_cast<float*>(smem);
    int* rp_all = reinterpret_cast<int*>(smem + NST * SE * 256);
    int* tq = rp_all + RPN * (TILE + 4);
    uint64_t* full = reinterpret_cast<uint64_t*>(tq + RPN);
    uint64_t* empty = full + NST;
    static_assert(RPN >= NST + 2, "rowptr slots must outlast the ring");

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t keep = pol_evict_last(), stream = pol_evict_first();
    const size_t plane = (size_t)a.N * a.B * H;

    if (warp == NCW) {
        // ================= producer warp =================
        uint32_t gs = 0;                                   // stages produced so far
        for (int k = 0;; ++k) {
            int tile = 0;
            if (lane == 0) tile = atomicAdd(a.counter, 1);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            const bool last = tile >= a.n_tiles;
            int* rp_s = rp_all + (k % RPN) * (TILE + 4);
            int b = 0, n0 = 0, nrows = 0;
            if (!last) {
                b = tile / a.blocks_per_trial;
                const int blk = tile - b * a.blocks_per_trial;
                n0 = blk * TILE; nrows = min(TILE, a.N - n0);
                for (int i = lane; i <= TILE; i += 32) rp_s[i] = a.rowptr[n0 + min(i, nrows)];
            }
            if (lane == 0) tq[k % RPN] = last ? -1 : tile;
            __syncwarp();
            int e0 = 0, etot = 0, nst_tile = 1;
            if (!last) { e0 = rp_s[0]; etot = rp_s[TILE] - e0; nst_tile = max(1, (etot + SE - 1) / SE); }
            const int row0 = b * a.N;
            const int half = lane >> 4, gl = lane & 15;           // lanes 0-15: stage 2i, lanes 16-31: stage 2i+1
            // index registers, fetched two iterations ahead (two explicit buffers: no dynamic register indexing)
            int idxA[4], idxB[4];
            auto load_idx = [&](int it, int (&d)[4]) {
                const int s = 2 * it + half;
                const int q = s * SE + 4 * gl;
#pragma unroll
                for (int u = 0; u < 4; ++u) d[u] = (q + u < etot) ? a.colidx[e0 + q + u] + row0 : -1;
            };
            auto issue = [&](int it, const int (&cur)[4]) {
                const int s = 2 * it + half;
                if (s < nst_tile) {
                    const uint32_t g = gs + s, st = g % NST;
                    mbar_wait(empty + st, ((g / NST) & 1u) ^ 1u);
                    const int n_ent = last ? 0 : min(SE, etot - s * SE);
                    const int ngrp = (n_ent + 3) >> 2;
                    if (gl == 0) mbar_arrive_expect_tx(full + st, (uint32_t)ngrp * 1024u);
                    __syncwarp(half ? 0xffff0000u : 0x0000ffffu);
                    if (gl < ngrp) {
                        const int r0 = cur[0];
                        const int r1 = cur[1] >= 0 ? cur[1] : r0, r2 = cur[2] >= 0 ? cur[2] : r0, r3 = cur[3] >= 0 ? cur[3] : r0;
                        tma_gather4(ring + (size_t)(st * SE + 4 * gl) * H, &tmX, 0, r0, r1, r2, r3, full + st, keep);
                    }
                }
                __syncwarp();
            };
            const int n_it = (nst_tile + 1) / 2;
            load_idx(0, idxA);
            if (n_it > 1) load_idx(1, idxB);
#pragma unroll 1
            for (int it = 0; it < n_it; it += 2) {
                int cur[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) cur[u] = idxA[u];
                if (it + 2 < n_it) load_idx(it + 2, idxA);
                issue(it, cur);
                if (it + 1 < n_it) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) cur[u] = idxB[u];
                    if (it + 3 < n_it) load_idx(it + 3, idxB);
                    issue(it + 1, cur);
                }
            }
            gs += nst_tile;
            if (last) break;
        }
    } else {
        // ================= consumer warps =================
        const int l = lane & 15, hsel = lane >> 4;
        uint32_t gs = 0, rel = 0;                          // stages consumed by tiles so far; next stage this warp releases
        auto release_to = [&](uint32_t upto) {             // warp-uniform
            // a stage may only be released after its "full" phase completed: arrivals of two uses of a slot never mix
            __syncwarp();
            for (uint32_t s = rel; s < upto; ++s) {
                mbar_wait(full + (s % NST), (s / NST) & 1u);
                if (lane == 0) mbar_arrive(empty + (s % NST));
            }
            rel = max(rel, upto);
        };
        for (int k = 0;; ++k) {
            mbar_wait(full + (gs % NST), (gs / NST) & 1u);        // first stage of the tile: rp_s / tq are published
            const int tile = tq[k % RPN];
            if (tile < 0) break;
            const int* rp_s = rp_all + (k % RPN) * (TILE + 4);
            const int b = tile / a.blocks_per_trial, blk = tile - b * a.blocks_per_trial;
            const int n0 = blk * TILE, nrows = min(TILE, a.N - n0);
            const int row0 = b * a.N;
            const int e0 = rp_s[0], etot = rp_s[TILE] - e0;
            const int nst_tile = max(1, (etot + SE - 1) / SE);
            float4 o0, o1, o2, o3;
            auto load_own = [&](int j) {
                const int rr = 2 * (warp + NCW * j) + hsel;
                const size_t off = (size_t)(row0 + n0 + min(rr, nrows - 1)) * H + 4 * l;
                o0 = ldg4_hint(a.X + off, keep);
                o1 = ldg4_hint(a.Y + off, stream); o2 = ldg4_hint(a.Y + plane + off, stream); o3 = ldg4_hint(a.Y + 2 * plane + off, stream);
            };
            if (a.skel) load_own(0);
#pragma unroll 1
            for (int j = 0; j < TILE / (2 * NCW); ++j) {
                const int p = warp + NCW * j;
                const int rr = 2 * p + hsel;
                const int qa = rp_s[rr] - e0, qb = rp_s[rr + 1] - e0;            // my row's entries [qa, qb)
                const int pa = rp_s[2 * p] - e0, pb = rp_s[2 * p + 2] - e0;      // the pair's entries (warp-uniform)
                float4 c0 = o0, c1 = o1, c2 = o2, c3 = o3;
                if (a.skel && j + 1 < TILE / (2 * NCW)) load_own(j + 1);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pb > pa) {
                    const int sA = pa / SE, sB = (pb - 1) / SE;
                    release_to(gs + sA);
                    for (int s = sA; s <= sB; ++s) {
                        const uint32_t g = gs + s, st = g % NST;
                        mbar_wait(full + st, (g / NST) & 1u);
                        const int lo = max(qa, s * SE), hi = min(qb, (s + 1) * SE);
                        const float* base = ring + ((ptrdiff_t)st * SE - (ptrdiff_t)s * SE) * H + 4 * l;
#pragma unroll 4
                        for (int q = lo; q < hi; ++q) {
                            const float4 v = *reinterpret_cast<const float4*>(base + (ptrdiff_t)q * H);
                            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                        }
                        if (s < sB) release_to(g + 1);
                    }
                }
                if (rr < nrows) {
                    const size_t off = (size_t)(row0 + n0 + rr) * H + 4 * l;
                    if (a.skel) {
                        stg4_hint(a.Yout + off, make_float4(c1.x + acc.x, c1.y + acc.y, c1.z, c1.w), stream);
                        stg4_hint(a.Yout + plane + off, make_float4(c2.x + c0.x, c2.y, c2.z + acc.z, c2.w), stream);
                        stg4_hint(a.Yout + 2 * plane + off, make_float4(c3.x, c3.y + c0.y, c3.z, c3.w + acc.w), stream);
                    }
                    stg4_hint(a.out + off, acc, stream);
                }
            }
            release_to(gs + nst_tile);
            gs += nst_tile;
        }
    }
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void ba_graph(int N, int m, std::vector<int>& rowptr, std::vector<int>& colidx) {
    std::mt19937 rng(0);
    std::vector<int> rep;                // every endpoint once per incident edge (preferential attachment)
    std::vector<std::vector<int>> adj(N);
    for (int i = 0; i < m; ++i) rep.push_back(i);
    for (int v = m; v < N; ++v) {
        std::vector<int> tg;
        while ((int)tg.size() < m) {
            int t = rep[rng() % rep.size()];
            if (std::find(tg.begin(), tg.end(), t) == tg.end()) tg.push_back(t);
        }
        for (int t : tg) { adj[v].push_back(t); adj[t].push_back(v); rep.push_back(t); rep.push_back(v); }
    }
    rowptr.assign(N + 1, 0);
    for (int v = 0; v < N; ++v) { std::sort(adj[v].begin(), adj[v].end()); rowptr[v + 1] = rowptr[v] + (int)adj[v].size(); }
    colidx.resize(rowptr[N]);
    for (int v = 0; v < N; ++v) std::copy(adj[v].begin(), adj[v].end(), colidx.begin() + rowptr[v]);
}

template <int NST>
static float run_ring(const Args& a, const CUtensorMap& tm, int sms) {
    const size_t smem = (size_t)NST * SE * 256 + RPN * (TILE + 4) * 4 + RPN * 4 + 2 * NST * 8 + 1024 + 64;
    CK(cudaFuncSetAttribute(gather_ring<NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaMemset(a.counter, 0, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    gather_ring<NST><<<sms, (NCW + 1) * 32, smem>>>(a, tm);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 64;
    const int nst = argc > 2 ? atoi(argv[2]) : 3;
    const int reps = argc > 3 ? atoi(argv[3]) : 5;
    const int N = 75879, m = 5;
    std::vector<int> rowptr, colidx;
    ba_graph(N, m, rowptr, colidx);
    const size_t M = (size_t)N * B;
    const double nnzB = (double)colidx.size() * B;
    int maxdeg = 0; for (int v = 0; v < N; ++v) maxdeg = std::max(maxdeg, rowptr[v + 1] - rowptr[v]);
    printf("BA N=%d m=%d nnz=%zu maxdeg=%d trials=%d rows=%zu\n", N, m, colidx.size(), maxdeg, B, M);

    int *d_rp, *d_ci, *d_cnt;
    float *d_X, *d_Y, *d_out, *d_out2, *d_Yout;
    CK(cudaMalloc(&d_rp, rowptr.size() * 4)); CK(cudaMalloc(&d_ci, colidx.size() * 4)); CK(cudaMalloc(&d_cnt, 4));
    CK(cudaMemcpy(d_rp, rowptr.data(), rowptr.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ci, colidx.data(), colidx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_X, M * H * 4)); CK(cudaMalloc(&d_out, M * H * 4)); CK(cudaMalloc(&d_out2, M * H * 4));
    CK(cudaMalloc(&d_Y, 3 * M * H * 4)); CK(cudaMalloc(&d_Yout, 3 * M * H * 4));
    {
        std::vector<float> h(M * H);
        std::mt19937 rng(1);
        for (auto& v : h) v = (float)(rng() & 0xffff) / 65536.0f;
        CK(cudaMemcpy(d_X, h.data(), M * H * 4, cudaMemcpyHostToDevice));
        for (int p = 0; p < 3; ++p) CK(cudaMemcpy(d_Y + p * M * H, h.data(), M * H * 4, cudaMemcpyHostToDevice));
    }
    // tensor map over X: [M rows][64 floats], box {64, 1} (tile::gather4 fetches 4 such rows)
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {H, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {H * 4};
    const cuuint32_t box[2] = {H, (cuuint32_t)(argc > 4 ? atoi(argv[4]) : 1)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_X, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 1; }

    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    Args a;
    a.rowptr = d_rp; a.colidx = d_ci; a.N = N; a.B = B; a.blocks_per_trial = (N + TILE - 1) / TILE;
    a.n_tiles = a.blocks_per_trial * B; a.X = d_X; a.Y = d_Y; a.Yout = d_Yout; a.counter = d_cnt;

    for (int skel = 0; skel < 2; ++skel) {
        a.skel = skel;
        const double bytes_alg = skel ? (double)M * 2060.0 : (double)M * 512.0;      // algorithmic bytes (perfect reuse of gathered rows)
        const double bytes_g = nnzB * 256.0;
        float best_ldg = 1e30f, best_ring = 1e30f;
        for (int r = 0; r < reps; ++r) {
            a.out = d_out;
            CK(cudaMemset(d_cnt, 0, 4));
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0));
            gather_ldg<8, 2><<<2 * sms, 512>>>(a);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            best_ldg = std::min(best_ldg, ms);
            a.out = d_out2;
            float mr = nst == 3 ? run_ring<3>(a, tm, sms) : nst == 4 ? run_ring<4>(a, tm, sms) : nst == 6 ? run_ring<6>(a, tm, sms) : run_ring<2>(a, tm, sms);
            best_ring = std::min(best_ring, mr);
        }
        {   // sweep: loads in flight per lane x resident CTAs per SM
            auto timeit = [&](auto kern, int grid, const char* name, int dsm = 0) {
                float best = 1e30f;
                CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dsm));
                for (int r = 0; r < 3; ++r) {
                    CK(cudaMemset(d_cnt, 0, 4));
                    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                    CK(cudaEventRecord(e0));
                    kern<<<grid, 512, dsm>>>(a);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    best = std::min(best, ms);
                }
                printf("  %s %-28s %.3f ms  %.3e rows/s\n", skel ? "skel" : "aggr", name, best, M / (best * 1e-3));
            };
            a.out = d_out;
            timeit(gather_ldg<4, 2>, 2 * sms, "batch 4, 2 CTA/SM");
            timeit(gather_ldg<8, 2>, 2 * sms, "batch 8, 2 CTA/SM");
            timeit(gather_ldg<8, 2>, 1 * sms, "batch 8, 1 CTA/SM");
            timeit(gather_ldg<12, 2>, 2 * sms, "batch 12, 2 CTA/SM");
            timeit(gather_ldg<16, 1>, 1 * sms, "batch 16, 1 CTA/SM (128 regs)");
            timeit(gather_ldg<4, 2>, 1 * sms, "batch 4, 1 CTA/SM");
            timeit(gather_ldg<8, 2>, 2 * sms, "b8 2CTA dynsmem 32K", 32768);
            timeit(gather_ldg<8, 2>, 2 * sms, "b8 2CTA dynsmem 64K", 65536);
            timeit(gather_ldg<8, 2>, 2 * sms, "b8 2CTA dynsmem 96K", 98304);
            timeit(gather_ldg<8, 2>, 2 * sms, "b8 2CTA dynsmem 110K", 112640);
            timeit(gather_ldg<8, 2, 0, 8>, 2 * sms, "b8 2CTA dynsmem 96K + 16 smem wf/row", 98304);
            timeit(gather_ldg<8, 2, 0, 16>, 2 * sms, "b8 2CTA dynsmem 96K + 32 smem wf/row", 98304);
            timeit(gather_ldg<8, 2, 0, 28>, 2 * sms, "b8 2CTA dynsmem 96K + 56 smem wf/row", 98304);
            timeit(gather_ldg<8, 2, 1>, 2 * sms, "b8 2CTA noalloc dynsmem 0", 0);
            timeit(gather_ldg<8, 2, 1>, 2 * sms, "b8 2CTA noalloc dynsmem 110K", 112640);
            CK(cudaMemset(d_cnt, 0, 4));
            gather_ldg<8, 2><<<2 * sms, 512>>>(a);     // restore d_out for the comparison
            CK(cudaDeviceSynchronize());
        }
        // compare
        std::vector<float> h1(1 << 20), h2(1 << 20);
        size_t bad = 0;
        for (size_t off = 0; off < M * H; off += (M * H / 7 / 4) * 4) {
            const size_t n = std::min<size_t>(1 << 20, M * H - off);
            CK(cudaMemcpy(h1.data(), d_out + off, n * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h2.data(), d_out2 + off, n * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; ++i) if (h1[i] != h2[i]) { if (bad < 5) printf("  mismatch at %zu: %g vs %g\n", off + i, h1[i], h2[i]); ++bad; }
        }
        printf("%s: LDG  %.3f ms  %.3e rows/s  gathered %.0f GB/s  algorithmic %.0f GB/s\n", skel ? "skeleton   " : "aggregation", best_ldg,
               M / (best_ldg * 1e-3), bytes_g / best_ldg * 1e-6, bytes_alg / best_ldg * 1e-6);
        printf("%s: RING %.3f ms  %.3e rows/s  gathered %.0f GB/s  algorithmic %.0f GB/s  (nst=%d)  mismatches=%zu\n", skel ? "skeleton   " : "aggregation",
               best_ring, M / (best_ring * 1e-3), bytes_g / best_ring * 1e-6, bytes_alg / best_ring * 1e-6, nst, bad);
    }
    return 0;
}
