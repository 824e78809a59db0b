// Probe: where do the rows of a cta_group::1, M=64 tcgen05.mma accumulator live in TMEM?
// A[r][0] = r + 1 (other k = 0), B[n][0] = 1  ->  D[r][n] = r + 1. Every warp dumps its 32 lanes x 16 columns.
#include <cstdio>
#include <cuda_runtime.h>
#include "../gn-ode-sir_b200/csrc/gnode_umma.cuh"
using namespace gnode;

__global__ void probe(float* out, int M) {
    __shared__ __align__(1024) unsigned char As[128 * 128];   // K-block of 32 fp32, 128 rows
    __shared__ __align__(1024) unsigned char Bs[16 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 128 * 32; i += blockDim.x) reinterpret_cast<float*>(As)[i] = 0.f;
    for (int i = tid; i < 16 * 32; i += blockDim.x) reinterpret_cast<float*>(Bs)[i] = 0.f;
    __syncthreads();
    if (tid < 128) *reinterpret_cast<float*>(As + sw_off(tid, 0)) = (float)(tid + 1);         // element (r, k=0)
    if (tid < 16) *reinterpret_cast<float*>(Bs + (tid << 7) + (((0 ^ (tid & 7))) << 4)) = 1.f;
    if (tid < 32) umma::tmem_alloc(&tslot, 32);
    if (tid == 0) umma::mbar_init(&bar, 1);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tslot;
    if (tid == 0) {
        const uint32_t idesc = umma::instr_desc_tf32(M, 16);
        umma::mma_tf32(tmem, umma::smem_desc(umma::smem_u32(As)), umma::smem_desc(umma::smem_u32(Bs)), idesc, 0);
        umma::mma_commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)((tid >> 5) * 32) << 16), v);
    for (int c = 0; c < 16; ++c) out[tid * 16 + c] = v[c];
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tmem, 32);
}

int main() {
    float* d; cudaMalloc(&d, 128 * 16 * 4);
    for (int M : {128, 64}) {
        cudaMemset(d, 0xff, 128 * 16 * 4);
        probe<<<1, 128>>>(d, M);
        cudaError_t e = cudaDeviceSynchronize();
        printf("M=%d: %s\n", M, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        float h[128 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        for (int lane = 0; lane < 128; ++lane) {
            if (lane % 16 == 0) printf("\n lanes %3d..: ", lane);
            printf("%g/%g ", h[lane * 16 + 0], h[lane * 16 + 5]);
        }
        printf("\n");
    }
    return 0;
}
