// libsplendor_b200.so - self-test of the tcgen05 / TMEM building blocks (spl_umma.cuh): one CTA computes
// C[128][N] = A[128][K] . B[N][K]^T (bf16 in, fp32 out) through shared-memory descriptors, tcgen05.mma, tcgen05.commit,
// an mbarrier and tcgen05.ld. The GPU tests compare it with a float64 product; the fused evaluator uses the same blocks.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "spl_internal.h"
#include "spl_umma.cuh"

namespace {

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                               float* __restrict__ C, int N, int K, int* __restrict__ err) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, kc = K / 8;
    unsigned char* sA = smem;
    unsigned char* sB = smem + 128 * K * 2;
    for (int i = tid; i < 128 * kc; i += 128) {
        const int r = i / kc, c = i - r * kc;
        *reinterpret_cast<uint4*>(sA + umma::chunk_off(r, c, kc)) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + c * 8);
    }
    for (int i = tid; i < N * kc; i += 128) {
        const int r = i / kc, c = i - r * kc;
        *reinterpret_cast<uint4*>(sB + umma::chunk_off(r, c, kc)) = *reinterpret_cast<const uint4*>(B + (size_t)r * K + c * 8);
    }
    umma::fence_smem_to_async();
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
    if (tid == 0) umma::mbar_init(&bar, 1);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = umma::instr_desc_bf16(128, N);
        for (int k = 0; k < K / 16; k++) {
            const uint64_t ad = umma::smem_desc(umma::smem_u32(sA) + k * 256, 128, kc * 128);
            const uint64_t bd = umma::smem_desc(umma::smem_u32(sB) + k * 256, 128, kc * 128);
            umma::mma_bf16_ss(tbase, ad, bd, idesc, k > 0);
        }
        umma::commit(&bar);
    }
    const bool ok = umma::mbar_wait(&bar, 0, 1u << 22);
    umma::fence_after_sync();
    if (!ok) {
        if (lane == 0) atomicExch(err, 1);
    } else {
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            umma::tmem_ld32(umma::tmem_addr(tbase, warp * 32, c0), v);
            for (int j = 0; j < 32; j++) C[(size_t)(warp * 32 + lane) * N + c0 + j] = v[j];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 256);
}

}   // namespace

extern "C" int spl_umma_selftest(spl_ctx* c, const void* a_bf16, const void* b_bf16, float* out, int n, int k, int* err_flag, void* stream) {
    if (!c) return spl_fail_(SPL_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (!a_bf16 || !b_bf16 || !out || !err_flag || n < 32 || n > 256 || n % 32 || k < 16 || k > 256 || k % 16)
        return spl_fail_(SPL_E_ARG, "spl_umma_selftest: bad argument");
    const int smem = (128 + n) * k * 2;
    CU(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)b_bf16, out, n, k, err_flag);
    CU(cudaGetLastError());
    return SPL_OK;
}
