// Probe: what does tcgen05.mma kind::tf32 do with the low 13 mantissa bits of a 32-bit operand element?
// A[r][0] = x_r (full fp32 mantissa), B[n][0] = 1  ->  D[r][n] = the value the tensor core used for x_r.
// Prints how many rows match trunc (bits & 0xFFFFE000), round-to-nearest-away (cvt.rna.tf32) or neither.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../gn-ode-sir_b200/csrc/gnode_umma.cuh"
using namespace gnode;

__global__ void probe(const float* xin, float* out) {
    __shared__ __align__(1024) unsigned char As[128 * 128];
    __shared__ __align__(1024) unsigned char Bs[16 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 128 * 32; i += blockDim.x) reinterpret_cast<float*>(As)[i] = 0.f;
    for (int i = tid; i < 16 * 32; i += blockDim.x) reinterpret_cast<float*>(Bs)[i] = 0.f;
    __syncthreads();
    if (tid < 128) *reinterpret_cast<float*>(As + sw_off(tid, 0)) = xin[tid];
    if (tid < 16) *reinterpret_cast<float*>(Bs + (tid << 7) + (((0 ^ (tid & 7))) << 4)) = 1.f;
    if (tid < 32) umma::tmem_alloc(&tslot, 32);
    if (tid == 0) umma::mbar_init(&bar, 1);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tslot;
    if (tid == 0) {
        const uint32_t idesc = umma::instr_desc_tf32(128, 16);
        umma::mma_tf32(tmem, umma::smem_desc(umma::smem_u32(As)), umma::smem_desc(umma::smem_u32(Bs)), idesc, 0);
        umma::mma_commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)((tid >> 5) * 32) << 16), v);
    out[tid] = v[0];
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tmem, 32);
}

int main() {
    float hx[128], hd[128];
    for (int r = 0; r < 128; ++r) {
        uint32_t low = (r < 8) ? (uint32_t[]){0x1FFF, 0x1000, 0x1001, 0x0FFF, 0x0001, 0x0000, 0x1800, 0x0800}[r] : (uint32_t)((r * 2654435761u) >> 19);
        uint32_t bits = 0x3F800000u | (((uint32_t)r & 15u) << 13) | (low & 0x1FFFu) | ((r & 16) ? 0x80000000u : 0u);
        if (r >= 64) bits = (bits & 0x807FFFFFu) | ((uint32_t)(100 + r) << 23);     // other exponents
        memcpy(&hx[r], &bits, 4);
    }
    float *dx, *dd; cudaMalloc(&dx, 512); cudaMalloc(&dd, 512);
    cudaMemcpy(dx, hx, 512, cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(dx, dd);
    cudaError_t e = cudaDeviceSynchronize();
    printf("probe: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(hd, dd, 512, cudaMemcpyDeviceToHost);
    int n_trunc = 0, n_rna = 0, n_other = 0;
    for (int r = 0; r < 128; ++r) {
        uint32_t xb, db; memcpy(&xb, &hx[r], 4); memcpy(&db, &hd[r], 4);
        const uint32_t tr = xb & 0xFFFFE000u;
        const uint32_t rn = (xb + 0x1000u) & 0xFFFFE000u;           // round to nearest, ties away (magnitude)
        const bool is_t = db == tr, is_r = db == rn;
        n_trunc += is_t; n_rna += is_r; n_other += (!is_t && !is_r);
        if (r < 12 || (!is_t && !is_r)) printf(" r=%3d x=%08x used=%08x trunc=%08x rna=%08x %s%s\n", r, xb, db, tr, rn, is_t ? "T" : "", is_r ? "R" : "");
    }
    printf("rows matching trunc: %d, rna: %d, neither: %d (of 128; rows whose low bits are 0 match both)\n", n_trunc, n_rna, n_other);
    return 0;
}
