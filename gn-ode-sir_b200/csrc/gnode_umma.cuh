// tcgen05 (5th-gen tensor core) helpers for the node-wise H x H transform:
//   dst = sigmoid(X W^T + b),  X: [128 rows x 64] fp32 tile, W: [64 x 64] fp32
// computed as a 3xTF32 split product so that the result is fp32-accurate
// (BASELINE.json north_star: "3xTF32 or fp32 so results stay within tolerance"):
//   X = Xhi + Xlo,  W = Whi + Wlo   (hi = rna_tf32(x), lo = rna_tf32(x - hi); |x - hi - lo| <= 2^-23 |x|)
//   X W^T = Xlo Wlo^T + Xlo Whi^T + Xhi Wlo^T + Xhi Whi^T   (all four terms: per-product error ~2^-22)
// Operands live in shared memory in the canonical UMMA K-major SWIZZLE_128B layout
// (two K-blocks of 32 fp32; rows of 128 B; 16-B chunks XORed with row&7), the fp32
// accumulator [128 lanes x 64 columns] lives in TMEM, tcgen05.mma is issued by one thread,
// completion is signalled through an mbarrier by tcgen05.commit, and the four "row warps"
// read the accumulator back with tcgen05.ld (lane == tile row) for the bias + sigmoid epilogue.
#pragma once
#include "gnode_common.cuh"

namespace gnode {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one warp, .sync.aligned)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    uint32_t spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) __trap();      // a lost completion must fault, never hang the GPU
    } while (!done);
}

// same, with a suspend-time hint. Measured on B200: the hint does not stop the polling (~40 issued instructions
// per row are spent here), but both alternatives were slower: a __nanosleep(64) back-off (-4 %, late wake-up) and
// one polling warp followed by a block barrier (equal).
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    uint32_t spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (!done && ++spins > (1u << 24)) __trap();      // a lost completion must fault, never hang the GPU
    } while (!done);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier among a subset of the CTA's warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- descriptors (cute/arch/mma_sm100_desc.hpp bit layout)
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: start>>4 [0,14) | LBO>>4 [16,30) (ignored for
// swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between 8-row groups | version=1 [46,48) | layout=2 [61,64)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor, kind::tf32: D=F32 [4,6)=1 | A=TF32 [7,10)=2 | B=TF32 [10,13)=2 | A,B K-major |
// N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 4 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- TF32 split: x = hi + lo + O(2^-23 |x|), hi = rna_tf32(x), lo = rna_tf32(x - hi); both have their
// low 13 mantissa bits zero, so the tensor core (which reads only the top 19 bits) sees them exactly.
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) { hi = tf32_rna(x); lo = tf32_rna(x - hi); }
__device__ __forceinline__ void tf32_split4(float4 x, float4& hi, float4& lo) {
    tf32_split(x.x, hi.x, lo.x); tf32_split(x.y, hi.y, lo.y); tf32_split(x.z, hi.z, lo.z); tf32_split(x.w, hi.w, lo.w);
}

// Truncation split (the pipelined step kernels): tcgen05.mma kind::tf32 reads only the top 19 bits of a 32-bit element
// (tools/umma_lowbits_probe.cu), so a raw fp32 word IS the operand hi = trunc_tf32(x): nothing is converted or written
// back. lo = rna_tf32(x - hi); x - hi has up to 13 significant bits, so |x - hi - lo| <= 2^-21 |x| (rna split: 2^-22).
// (cvt.rna.tf32.f32 is emulated on sm_100a: an Inf/NaN test, a predicated add of half an ulp, a mask. The residual is
// finite and tiny, so the add and the mask are applied directly: the same bits, one instruction less per element.)
__device__ __forceinline__ float tf32_trunc_lo(float x) {
    const float d = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    return __uint_as_float((__float_as_uint(d) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float4 tf32_trunc_lo4(float4 x) {
    return make_float4(tf32_trunc_lo(x.x), tf32_trunc_lo(x.y), tf32_trunc_lo(x.z), tf32_trunc_lo(x.w));
}

// byte offset of chunk c4 of row n in the 64-row B operand (K-block stride 8 KB)
__device__ __forceinline__ int swb_off(int n, int c4) { return ((c4 >> 3) << 13) + (n << 7) + (((c4 & 7) ^ (n & 7)) << 4); }

constexpr int TMEM_COLS = 64;

struct Ctx {
    uint32_t tmem;       // TMEM base address of the [128 x 64] fp32 accumulator
    uint64_t* bar;       // mbarrier signalled by tcgen05.commit
    uint32_t phase;      // parity of the next completion
    uint32_t whi, wlo;   // shared-space addresses of the W operand tiles (hi, lo)
};

// W [64 x 64] fp32 (global, [out][in]) -> Whi / Wlo operand tiles. All threads; caller syncs.
__device__ __forceinline__ void prepare_weights(const float* __restrict__ W, unsigned char* Whi, unsigned char* Wlo,
                                                int tid, int nthreads) {
    for (int idx = tid; idx < H * CHUNKS; idx += nthreads) {
        const int n = idx >> 4, c4 = idx & 15;
        float4 hi, lo;
        tf32_split4(ldg4(W + n * H + 4 * c4), hi, lo);
        sts4(Whi, swb_off(n, c4), hi);
        sts4(Wlo, swb_off(n, c4), lo);
    }
    fence_proxy_async();
}

// One thread: D[128x64] = Xlo Wlo^T + Xlo Whi^T + Xhi Wlo^T + Xhi Whi^T (small terms first), then commit.
// 4 passes x 8 K-steps of tcgen05.mma kind::tf32 (M128 N64 K8).
__device__ __forceinline__ void issue_split_gemm_to(uint32_t tmem, uint64_t* bar, uint32_t whi, uint32_t wlo,
                                                    uint32_t xhi, uint32_t xlo) {
    fence_after_sync();
    constexpr uint32_t idesc = instr_desc_tf32(TILE, H);
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const uint32_t abase = (pass < 2) ? xlo : xhi;
        const uint32_t bbase = (pass == 0 || pass == 2) ? wlo : whi;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t aoff = ((k >> 2) << 14) + ((k & 3) << 5);
            const uint32_t boff = ((k >> 2) << 13) + ((k & 3) << 5);
            mma_tf32(tmem, smem_desc(abase + aoff), smem_desc(bbase + boff), idesc, acc);
            acc = 1;
        }
    }
    mma_commit(bar);
}

// ---- N = 80 variant: the B operand stacks linear.weight (rows 0..63) on linear3.weight (rows 64..67, rows 68..79
// zero), so the accumulator columns 64..67 hold the decoder's hidden pre-activations of the same rows
// (ode_nn_ngraph_sim.py:172-176) at no extra operand traffic. K-block stride of the 80-row operand: 80 * 128 B.
constexpr int NB80 = 80;
constexpr int WB80_KBLOCK = NB80 * 128;            // 10240 B (1024-B aligned)
constexpr int WB80_BYTES = 2 * WB80_KBLOCK;        // 20480 B per operand (hi or lo)
__device__ __forceinline__ int swb80_off(int n, int c4) { return (c4 >> 3) * WB80_KBLOCK + (n << 7) + (((c4 & 7) ^ (n & 7)) << 4); }

__device__ __forceinline__ void prepare_weights80(const float* __restrict__ W, const float* __restrict__ W3, unsigned char* Whi,
                                                  unsigned char* Wlo, int tid, int nthreads) {
    for (int idx = tid; idx < NB80 * CHUNKS; idx += nthreads) {
        const int n = idx >> 4, c4 = idx & 15;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
        if (n < H) x = ldg4(W + n * H + 4 * c4);
        else if (n < H + 4) x = ldg4(W3 + (n - H) * H + 4 * c4);
        tf32_split4(x, hi, lo);
        sts4(Whi, swb80_off(n, c4), hi);
        sts4(Wlo, swb80_off(n, c4), lo);
    }
    fence_proxy_async();
}

// One thread: D[TR x 80] (TMEM columns tmem .. tmem+79) = X [W; W3]^T as the 4-term split product, then commit.
// TR = UMMA M = 128 or 64; the A operand's K-block stride is TR * 128 B.
template <int TR, int P0 = 0>      // P0 > 0: timing ablations only (passes skipped, wrong numerics)
__device__ __forceinline__ void issue_split_gemm80(uint32_t tmem, uint64_t* bar, uint32_t whi, uint32_t wlo, uint32_t xhi, uint32_t xlo) {
    fence_after_sync();
    constexpr uint32_t idesc = instr_desc_tf32(TR, NB80);
    uint32_t acc = 0;
#pragma unroll
    for (int pass = P0; pass < 4; ++pass) {
        const uint32_t abase = (pass < 2) ? xlo : xhi;
        const uint32_t bbase = (pass == 0 || pass == 2) ? wlo : whi;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t aoff = (uint32_t)(k >> 2) * (TR * 128) + ((k & 3) << 5);
            const uint32_t boff = (uint32_t)(k >> 2) * WB80_KBLOCK + ((k & 3) << 5);
            mma_tf32(tmem, smem_desc(abase + aoff), smem_desc(bbase + boff), idesc, acc);
            acc = 1;
        }
    }
    mma_commit(bar);
}

// ---- N = 160 variant (the pipelined step kernels): the hi and the lo part of [W; W3] are stacked along N (rows 0..79 hi,
// rows 80..159 lo of ONE B operand), so a pass over an A tile produces A Whi^T in accumulator columns 0..79 and A Wlo^T
// in columns 80..159: the 4-term split product takes 16 MMAs of M128 N160 K8 (X lo pass, then X hi pass) instead of 32
// of N80, and every A tile is read from shared memory once per GEMM instead of twice (144 KB of operand reads per GEMM
// instead of 208 KB -- the tensor core fetches its operands over the same shared-memory port the LSU uses, and
// removing the MMAs altogether is worth 17 % of the step, profiles/r2n_ab_ablations.log). The epilogue adds the two
// column blocks (small part first). K-block stride of the 160-row operand: 160 * 128 B.
constexpr int NB160 = 160;
constexpr int WB160_KBLOCK = NB160 * 128;          // 20480 B (1024-B aligned)
constexpr int WB160_BYTES = 2 * WB160_KBLOCK;      // 40960 B = the hi and lo operands of the N = 80 variant together
__device__ __forceinline__ int swb160_off(int n, int c4) { return (c4 >> 3) * WB160_KBLOCK + (n << 7) + (((c4 & 7) ^ (n & 7)) << 4); }

__device__ __forceinline__ void prepare_weights160(const float* __restrict__ W, const float* __restrict__ W3, unsigned char* Wop,
                                                   int tid, int nthreads) {
    for (int idx = tid; idx < NB80 * CHUNKS; idx += nthreads) {
        const int n = idx >> 4, c4 = idx & 15;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
        if (n < H) x = ldg4(W + n * H + 4 * c4);
        else if (n < H + 4) x = ldg4(W3 + (n - H) * H + 4 * c4);
        tf32_split4(x, hi, lo);
        sts4(Wop, swb160_off(n, c4), hi);
        sts4(Wop, swb160_off(NB80 + n, c4), lo);
    }
    fence_proxy_async();
}

// One thread: one pass of 8 MMAs (M = TR, N = 160, K = 8) of the A tile at `abase` over the stacked operand; `fresh` = the
// first MMA overwrites the accumulator.
template <int TR>
__device__ __forceinline__ void issue_pass160(uint32_t tmem, uint32_t wop, uint32_t abase, bool fresh) {
    constexpr uint32_t idesc = instr_desc_tf32(TR, NB160);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t aoff = (uint32_t)(k >> 2) * (TR * 128) + ((k & 3) << 5);
        const uint32_t boff = (uint32_t)(k >> 2) * WB160_KBLOCK + ((k & 3) << 5);
        mma_tf32(tmem, smem_desc(abase + aoff), smem_desc(wop + boff), idesc, (fresh && k == 0) ? 0u : 1u);
    }
}
// One thread: TMEM columns tmem .. tmem+79 = X Whi^T, tmem+80 .. tmem+159 = X Wlo^T with X = Xlo + Xhi (small pass first),
// then commit. P0 > 0: timing ablation only (lo pass skipped, wrong numerics). (Issuing the hi pass of GEMM1 early, while
// the threads still compute the lo operand, was measured: no gain, and big-before-small accumulation costs accuracy.)
template <int TR, int P0 = 0>
__device__ __forceinline__ void issue_split_gemm160(uint32_t tmem, uint64_t* bar, uint32_t wop, uint32_t xhi, uint32_t xlo) {
    fence_after_sync();
    if (P0 == 0) issue_pass160<TR>(tmem, wop, xlo, true);
    issue_pass160<TR>(tmem, wop, xhi, P0 != 0);
    mma_commit(bar);
}

// A/B variant: the lo x lo term dropped (3xTF32). X hi pass over the N = 160 operand first (its first MMA overwrites all
// 160 columns), then the X lo pass over the hi rows only (N = 80, columns 0..79).
template <int TR>
__device__ __forceinline__ void issue_split_gemm160_3term(uint32_t tmem, uint64_t* bar, uint32_t wop, uint32_t xhi, uint32_t xlo) {
    fence_after_sync();
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const uint32_t abase = pass == 0 ? xhi : xlo;
        const uint32_t idesc = pass == 0 ? instr_desc_tf32(TR, NB160) : instr_desc_tf32(TR, NB80);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t aoff = (uint32_t)(k >> 2) * (TR * 128) + ((k & 3) << 5);
            const uint32_t boff = (uint32_t)(k >> 2) * WB160_KBLOCK + ((k & 3) << 5);
            mma_tf32(tmem, smem_desc(abase + aoff), smem_desc(wop + boff), idesc, acc);
            acc = 1;
        }
    }
    mma_commit(bar);
}

// 16 (4) accumulator columns of this thread's TMEM lane, summed over the two column blocks of the N = 160 product:
// v = (X Wlo^T)[c] + (X Whi^T)[c]
__device__ __forceinline__ void tmem_ld16_sum(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16], s[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]),
                   "=r"(s[8]), "=r"(s[9]), "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15])
                 : "r"(taddr + (uint32_t)NB80));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __fadd_rn(__uint_as_float(s[i]), __uint_as_float(r[i]));
}
// the same, summed on pairs: v2[i] = {s[2i] + r[2i], s[2i+1] + r[2i+1]}
__device__ __forceinline__ void tmem_ld16_sum2(uint32_t taddr, f32x2 (&v2)[8]) {
    uint32_t r[16], s[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]),
                   "=r"(s[8]), "=r"(s[9]), "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15])
                 : "r"(taddr + (uint32_t)NB80));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i)
        v2[i] = add2(pack2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])));
}
__device__ __forceinline__ void tmem_ld4_sum(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4], s[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]) : "r"(taddr + (uint32_t)NB80));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __fadd_rn(__uint_as_float(s[i]), __uint_as_float(r[i]));
}

__device__ __forceinline__ void issue_split_gemm(const Ctx& cx, uint32_t xhi, uint32_t xlo) {
    fence_after_sync();
    constexpr uint32_t idesc = instr_desc_tf32(TILE, H);
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const uint32_t abase = (pass < 2) ? xlo : xhi;
        const uint32_t bbase = (pass == 0 || pass == 2) ? cx.wlo : cx.whi;
#pragma unroll
        for (int k = 0; k < 8; ++k) {                                  // 8 K-steps of 8 tf32 (32 B)
            const uint32_t aoff = ((k >> 2) << 14) + ((k & 3) << 5);
            const uint32_t boff = ((k >> 2) << 13) + ((k & 3) << 5);
            mma_tf32(cx.tmem, smem_desc(abase + aoff), smem_desc(bbase + boff), idesc, acc);
            acc = 1;
        }
    }
    mma_commit(cx.bar);
}

// All 16 warps: wait for the accumulator, then warp w converts the [32 lanes (w&3)] x [16 columns (w>>2)]
// block: dst = sigmoid(acc + b) written thread-per-row into the swizzled tile dst.
template <bool FAST>
__device__ __forceinline__ void epilogue_sigmoid(Ctx& cx, unsigned char* dst, const float* bs, int warp, int lane) {
    mbar_wait(cx.bar, cx.phase);
    fence_after_sync();
    const int q = warp & 3, cq = warp >> 2;
    float v[16];
    tmem_ld16(cx.tmem + ((uint32_t)(q * 32) << 16) + 16 * cq, v);
    const int row = q * 32 + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 bb = *reinterpret_cast<const float4*>(bs + 16 * cq + 4 * j);
        float4 o;
        o.x = sigmoid_t<FAST>(v[4 * j + 0] + bb.x); o.y = sigmoid_t<FAST>(v[4 * j + 1] + bb.y);
        o.z = sigmoid_t<FAST>(v[4 * j + 2] + bb.z); o.w = sigmoid_t<FAST>(v[4 * j + 3] + bb.w);
        sts4(dst, sw_off(row, 4 * cq + j), o);
    }
    fence_before_sync();
    cx.phase ^= 1;
}

// dst(Ls) = sigmoid(Xs W^T + b). Xs: fp32 tile, replaced in place by its tf32 hi part; Ls: scratch for the
// lo part, then the result. Called by all NTHREADS threads. Ends WITHOUT a block barrier: the caller must
// __syncthreads() before other warps read Ls.
template <bool FAST>
__device__ __forceinline__ void gemm_sigmoid_tc(Ctx& cx, unsigned char* Xs, unsigned char* Ls, const float* bs, int tid) {
    for (int idx = tid; idx < TILE * CHUNKS; idx += NTHREADS) {
        const int off = sw_off(idx >> 4, idx & 15);
        float4 hi, lo;
        tf32_split4(lds4(Xs, off), hi, lo);
        sts4(Xs, off, hi);
        sts4(Ls, off, lo);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) issue_split_gemm(cx, smem_u32(Xs), smem_u32(Ls));
    epilogue_sigmoid<FAST>(cx, Ls, bs, tid >> 5, tid & 31);
}

}  // namespace umma
}  // namespace gnode
