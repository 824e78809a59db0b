// Probe: D[h][j] = sum_r G[r][h] X[r][j] (the weight gradient of the reverse sweep) as tcgen05.mma kind::tf32 with BOTH
// operands MN-major, read straight from the K-major SWIZZLE_128B tiles the backward already holds ([128 rows r][64] fp32,
// two 32-column blocks of 16 KB, 16-B chunks XORed with r & 7): the same bytes are the canonical MN-major SW128 layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-B units with LBO = 16 KB (next 32-column block) and SBO = 1 KB (next 8 rows).
// M = 64 (h), N = 64 (j), K = 8 rows per instruction, 16 instructions per tile. Prints which (LBO, SBO) variant and
// which TMEM lane mapping reproduce the CPU result.
// RESULT on B200 (profiles/r2k_umma_mn_probe.log): both variants complete without an error and leave an all-zero
// accumulator. kind::tf32 accepts MN-major operands only in the SWIZZLE_128B_BASE32B layout (32-byte swizzle granules,
// layout type 1; cutlass/gemm/collective/builders/sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout"), which the K-major SWIZZLE_128B tile of the state VJP is not: the two GEMMs cannot share a tile.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../gn-ode-sir_b200/csrc/gnode_umma.cuh"
using namespace gnode;

__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void probe(const float* G, const float* X, float* out, int variant) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = raw + ((1024u - (umma::smem_u32(raw) & 1023u)) & 1023u);
    unsigned char* Gs = smem;
    unsigned char* Xs = smem + 32768;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 65536 + 8);
    const int tid = threadIdx.x;
    for (int idx = tid; idx < 128 * 16; idx += blockDim.x) {
        const int r = idx >> 4, c4 = idx & 15;
        sts4(Gs, sw_off(r, c4), ldg4(G + r * 64 + 4 * c4));
        sts4(Xs, sw_off(r, c4), ldg4(X + r * 64 + 4 * c4));
    }
    if (tid < 32) umma::tmem_alloc(tslot, 64);
    if (tid == 0) umma::mbar_init(bar, 1);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tslot;
    if (tid == 0) {
        // a_major (bit 15) = b_major (bit 16) = MN
        const uint32_t idesc = umma::instr_desc_tf32(64, 64) | (1u << 15) | (1u << 16);
        const uint32_t lbo = variant == 0 ? 16384u : 1024u, sbo = variant == 0 ? 1024u : 16384u;
        for (int k = 0; k < 16; ++k)
            umma::mma_tf32(tmem, smem_desc_mn(umma::smem_u32(Gs) + k * 1024, lbo, sbo),
                           smem_desc_mn(umma::smem_u32(Xs) + k * 1024, lbo, sbo), idesc, k > 0 ? 1u : 0u);
        umma::mma_commit(bar);
    }
    umma::mbar_wait(bar, 0);
    umma::fence_after_sync();
    const int warp = tid >> 5, lane = tid & 31;
    for (int cb = 0; cb < 4; ++cb) {
        float v[16];
        umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 16 * cb, v);
        for (int c = 0; c < 16; ++c) out[(warp * 32 + lane) * 64 + 16 * cb + c] = v[c];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tmem, 64);
}

int main() {
    static float G[128 * 64], X[128 * 64], D[64 * 64], out[128 * 64];
    srand(1);
    for (int i = 0; i < 128 * 64; ++i) { G[i] = (float)(rand() % 15 - 7); X[i] = (float)(rand() % 15 - 7); }
    for (int h = 0; h < 64; ++h)
        for (int j = 0; j < 64; ++j) {
            float s = 0.f;
            for (int r = 0; r < 128; ++r) s += G[r * 64 + h] * X[r * 64 + j];
            D[h * 64 + j] = s;
        }
    float *dG, *dX, *dO;
    cudaMalloc(&dG, sizeof(G)); cudaMalloc(&dX, sizeof(X)); cudaMalloc(&dO, sizeof(out));
    cudaMemcpy(dG, G, sizeof(G), cudaMemcpyHostToDevice); cudaMemcpy(dX, X, sizeof(X), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 2048);
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(dO, 0xff, sizeof(out));
        probe<<<1, 128, 65536 + 2048>>>(dG, dX, dO, variant);
        cudaError_t e = cudaDeviceSynchronize();
        printf("variant %d (LBO %d, SBO %d): %s\n", variant, variant == 0 ? 16384 : 1024, variant == 0 ? 1024 : 16384, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(out, dO, sizeof(out), cudaMemcpyDeviceToHost);
        // lane mapping hypotheses: A: lane = h (0..63); B: lane = (h % 16) + 32 * (h / 16); C: lane = (h % 32) + 64 * (h / 32)
        const char* names[3] = {"lane = h", "lane = h%16 + 32*(h/16)", "lane = h%32 + 64*(h/32)"};
        for (int hyp = 0; hyp < 3; ++hyp) {
            int bad = 0, badT = 0;
            for (int h = 0; h < 64; ++h)
                for (int j = 0; j < 64; ++j) {
                    const int lane = hyp == 0 ? h : (hyp == 1 ? (h % 16) + 32 * (h / 16) : (h % 32) + 64 * (h / 32));
                    if (out[lane * 64 + j] != D[h * 64 + j]) ++bad;
                    if (out[lane * 64 + j] != D[j * 64 + h]) ++badT;
                }
            printf("  %-28s mismatches: D %d, D^T %d of 4096\n", names[hyp], bad, badT);
        }
        printf("  lane 0: %g %g %g %g | want D[0][0..3] %g %g %g %g | lane 16: %g lane 32: %g (D[16][0] %g D[32][0] %g)\n", out[0], out[1], out[2],
               out[3], D[0], D[1], D[2], D[3], out[16 * 64], out[32 * 64], D[16 * 64], D[32 * 64]);
    }
    return 0;
}
