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

// the same product with B held MN-major in shared memory (the layout the transposed evaluator's epilogues write): element (n, k) at
// (n / 8) * stride_n8 + (k / 8) * stride_k8 + (k % 8) * 16 + (n % 8) * 2; the descriptor's LBO / SBO come from the caller
__global__ void __launch_bounds__(128, 1) umma_selftest_mn_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                                  float* __restrict__ C, int N, int K, int stride_k8, int stride_n8, int lbo, int sbo,
                                                                  int* __restrict__ err) {
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
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i - n * K;
        *reinterpret_cast<__nv_bfloat16*>(sB + (n >> 3) * stride_n8 + (k >> 3) * stride_k8 + (k & 7) * 16 + (n & 7) * 2) = B[i];
    }
    umma::fence_smem_to_async();
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
    if (tid == 0) umma::mbar_init(&bar, 1);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = umma::instr_desc_bf16_bmn(128, N);
        for (int k = 0; k < K / 16; k++) {
            const uint64_t ad = umma::smem_desc(umma::smem_u32(sA) + k * 256, 128, kc * 128);
            const uint64_t bd = umma::smem_desc(umma::smem_u32(sB) + k * 2 * stride_k8, lbo, sbo);
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

__device__ long long g_sink;
// diagnostics: SM cycles of `reps` x `ksteps` MMAs (M = 128, N = n, K = 16 each) on zero operands, from the first issue to the arrival
// of the commit; b_mn: B operand MN-major (LBO 128, SBO = sbo bytes) instead of K-major (LBO 128, SBO = ksteps * 256)
__global__ void __launch_bounds__(128, 1) umma_cycles_kernel(int n, int ksteps, int b_mn_in, int sbo, int reps, long long* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar, bar2, bar3;
    const int b_mn = b_mn_in & 1;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    umma::fence_smem_to_async();
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
    if (tid == 0) {
        umma::mbar_init(&bar, 1); umma::mbar_init(&bar2, 1); umma::mbar_init(&bar3, 1);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(&bar3)) : "memory");   // phase 0 of bar3 is complete
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if ((b_mn_in & 12) && warp == 0) {
        // lean issue: descriptors built once, groups of 4 k-steps unrolled with immediate offsets; bit 2: one thread, bit 3: the whole
        // warp runs the loop and an elected lane issues (uniform registers); bit 1: a commit after every group
        const uint32_t idesc = b_mn ? umma::instr_desc_bf16_bmn(128, n) : umma::instr_desc_bf16(128, n);
        const uint32_t sA = umma::smem_u32(smem), sB = sA + 32 * 1024;
        const uint64_t ad0 = umma::smem_desc(sA, 128, 4 * 256), bd0 = umma::smem_desc(sB, 128, b_mn ? sbo : 4 * 256);
        const bool one = (b_mn_in & 4) != 0, com = (b_mn_in & 2) != 0;
        uint32_t el = 0;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(el));
        const bool me = one ? tid == 0 : el != 0u;
        const int groups = reps * ksteps / 4;
        const long long t0 = clock64();
        if (!one || tid == 0) {
            for (int g = 0; g < groups; g++) {
                if (me) {
#pragma unroll
                    for (int k = 0; k < 4; k++) umma::mma_bf16_ss(tbase, ad0 + (uint64_t)(16 * k), bd0 + (uint64_t)(16 * k), idesc, k > 0 || g > 0);
                    if (com) umma::commit(&bar2);
                    if (b_mn_in & 16) umma::mbar_wait(&bar3, 0u, 1u << 20);      // a wait on a phase that completed long ago
                    if (b_mn_in & 32) umma::fence_after_sync();
                    if (b_mn_in & 64) g_sink = clock64();
                }
                if (!one) __syncwarp();
            }
        }
        const long long t1 = clock64();
        if (me) {
            umma::commit(&bar);
            umma::mbar_wait(&bar, 0, 1u << 24);
            const long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
    } else if (tid == 0) {
        const uint32_t idesc = b_mn ? umma::instr_desc_bf16_bmn(128, n) : umma::instr_desc_bf16(128, n);
        const uint32_t sA = umma::smem_u32(smem), sB = sA + 32 * 1024;
        const long long t0 = clock64();
        for (int r = 0; r < reps; r++)
            for (int k = 0; k < ksteps; k++) {
                const uint64_t ad = umma::smem_desc(sA + k * 256, 128, ksteps * 256);
                const uint64_t bd = umma::smem_desc(sB + k * 256, 128, b_mn ? sbo : ksteps * 256);
                umma::mma_bf16_ss(tbase, ad, bd, idesc, k > 0 || r > 0);
                if ((b_mn_in & 2) && k == ksteps - 1) umma::commit(&bar2);     // a commit after every group of `ksteps` MMAs (nobody waits for it)
            }
        const long long t1 = clock64();
        umma::commit(&bar);
        umma::mbar_wait(&bar, 0, 1u << 24);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 256);
}

// diagnostics: how fast one SM can stream [tile_bytes] tiles of an L2-resident buffer into a ring of `depth` shared-memory slots while all
// CTAs of the grid do the same. mode 0: cp.async.bulk by one thread (mbarrier complete_tx); mode 1: 16-byte cp.async by the 32 lanes of
// one warp (commit groups). out[cta] = SM cycles for `tiles` tiles
__global__ void __launch_bounds__(256, 1) stream_cycles_kernel(const unsigned char* __restrict__ src, int src_tiles, int tile_bytes, int depth, int tiles,
                                                               int mode, long long* __restrict__ out) {
    // mode = kind + 16 * (issuing warps - 1) + 256 * (pieces per tile - 1); kind 0: cp.async.bulk, 1: cp.async 16 B by the lanes of the warp
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[16];
    __shared__ long long t_end[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kind = mode & 15, nw = ((mode >> 4) & 15) + 1, pieces = (mode >> 8) + 1;
    if (tid == 0) for (int i = 0; i < 16; i++) umma::mbar_init(&full[i], 1);
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nw) {
        // warp w streams tiles w, w + nw, ... through slots w, w + nw, ... of the ring (depth / nw slots each)
        const int my_depth = depth / nw;
        if (kind == 0) {
            if (lane == 0) {
                int issued = 0;
                for (int t = warp; t < tiles + my_depth * nw; t += nw, issued++) {
                    const int sl = warp + nw * (issued % my_depth);
                    if (issued >= my_depth) umma::mbar_wait(&full[sl], (uint32_t)(((issued - my_depth) / my_depth) & 1));
                    if (t < tiles) {
                        umma::mbar_expect(&full[sl], (uint32_t)tile_bytes);
                        const int pb = tile_bytes / pieces;
                        for (int p = 0; p < pieces; p++)
                            umma::bulk_load(smem + (size_t)sl * tile_bytes + p * pb, src + (size_t)(t % src_tiles) * tile_bytes + p * pb, (uint32_t)pb, &full[sl]);
                    }
                }
            }
        } else if (kind == 2) {
            // every lane < my_depth of the warp owns ONE slot and streams tiles through it on its own (is the ~840-cycle service time of a
            // bulk copy a per-thread or a per-warp property?)
            if (lane < my_depth) {
                const int sl = warp + nw * lane;
                int n_done = 0;
                for (int t = warp + nw * lane; t < tiles; t += nw * my_depth, n_done++) {
                    umma::mbar_expect(&full[sl], (uint32_t)tile_bytes);
                    umma::bulk_load(smem + (size_t)sl * tile_bytes, src + (size_t)(t % src_tiles) * tile_bytes, (uint32_t)tile_bytes, &full[sl]);
                    umma::mbar_wait(&full[sl], (uint32_t)(n_done & 1));
                }
            }
        } else {
            int issued = 0;
            for (int t = warp; t < tiles + (my_depth - 1) * nw; t += nw, issued++) {
                if (t < tiles) {
                    const unsigned char* s_ = src + (size_t)(t % src_tiles) * tile_bytes;
                    unsigned char* d = smem + (size_t)(warp + nw * (issued % my_depth)) * tile_bytes;
                    for (int i = lane * 16; i < tile_bytes; i += 512)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(umma::smem_u32(d + i)), "l"(s_ + i) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (issued >= my_depth - 1) {
                    switch (my_depth - 1) {
                        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
                        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
                        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
                        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
                        case 7: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
                        default: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
                    }
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) t_end[warp] = clock64();
    }
    __syncthreads();
    if (tid == 0) {
        long long e = 0;
        for (int w = 0; w < nw; w++) e = t_end[w] > e ? t_end[w] : e;
        out[blockIdx.x] = e - t0;
    }
}

}   // namespace

extern "C" int spl_umma_stream_cycles(spl_ctx* c, const void* src, int src_tiles, int tile_bytes, int depth, int tiles, int mode, int grid,
                                      long long* out, void* stream) {
    if (!c) return spl_fail_(SPL_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (!src || !out || src_tiles < 1 || tile_bytes < 512 || (tile_bytes & 511) || depth < 1 || depth > 12 || (size_t)depth * tile_bytes > 200 * 1024 ||
        tiles < 1 || grid < 1)
        return spl_fail_(SPL_E_ARG, "spl_umma_stream_cycles: bad argument");
    CU(cudaFuncSetAttribute(stream_cycles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    stream_cycles_kernel<<<grid, 256, 200 * 1024, (cudaStream_t)stream>>>((const unsigned char*)src, src_tiles, tile_bytes, depth, tiles, mode, out);
    CU(cudaGetLastError());
    return SPL_OK;
}

extern "C" int spl_umma_mma_cycles(spl_ctx* c, int n, int ksteps, int b_mn, int sbo, int reps, long long* out2, void* stream) {
    if (!c) return spl_fail_(SPL_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (!out2 || n < 16 || n > 256 || n % 16 || ksteps < 1 || ksteps > 8 || reps < 1 || sbo < 0 || (sbo & 15) || (size_t)(n / 8) * ((b_mn & 1) ? sbo : ksteps * 256) > 128 * 1024)
        return spl_fail_(SPL_E_ARG, "spl_umma_mma_cycles: bad argument");
    CU(cudaFuncSetAttribute(umma_cycles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    umma_cycles_kernel<<<1, 128, 160 * 1024, (cudaStream_t)stream>>>(n, ksteps, b_mn, sbo, reps, out2);
    CU(cudaGetLastError());
    return SPL_OK;
}

extern "C" int spl_umma_selftest_mn(spl_ctx* c, const void* a_bf16, const void* b_bf16, float* out, int n, int k, int stride_k8, int stride_n8,
                                    int lbo, int sbo, int* err_flag, void* stream) {
    if (!c) return spl_fail_(SPL_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (!a_bf16 || !b_bf16 || !out || !err_flag || n < 32 || n > 256 || n % 32 || k < 16 || k > 256 || k % 16 || stride_k8 < 128 || stride_n8 < 128 ||
        ((stride_k8 | stride_n8 | lbo | sbo) & 15))
        return spl_fail_(SPL_E_ARG, "spl_umma_selftest_mn: bad argument");
    const size_t need = (size_t)128 * k * 2 + (size_t)(n / 8 - 1) * stride_n8 + (size_t)(k / 8 - 1) * stride_k8 + 128;
    if (need > 200 * 1024) return spl_fail_(SPL_E_ARG, "spl_umma_selftest_mn: strides too large");
    CU(cudaFuncSetAttribute(umma_selftest_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    umma_selftest_mn_kernel<<<1, 128, need, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)b_bf16, out, n, k, stride_k8,
                                                                    stride_n8, lbo, sbo, err_flag);
    CU(cudaGetLastError());
    return SPL_OK;
}

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
