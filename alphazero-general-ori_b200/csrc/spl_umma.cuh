// tcgen05 / TMEM building blocks for sm_100a (inline PTX), used by the fused leaf evaluator (spl_nnet.cu).
//
// Operand layout (both A [M rows][K] and B [N rows][K], bf16, K-major, no swizzle = the "interleave" canonical layout):
// 8 x 8 core matrices of 128 contiguous bytes (8 rows x 16 bytes); element (r, k) of an operand with KC = K / 8 core
// matrices per row group sits at byte
//     (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2,      LBO = 128, SBO = KC * 128
// i.e. a row group (8 rows) is one contiguous run of KC core matrices. One tcgen05.mma consumes K = 16 (two core matrices
// along K); the k-th instruction of a tile starts 256 bytes further.
// Accumulators: D [M = 128][N] fp32 in TMEM, row r = TMEM lane r, column c = TMEM column base + c. Warp w of a CTA may only
// touch lanes 32 (w % 4) .. +31, so a thread of warp w reads row 32 (w % 4) + lane.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46),
// version = 1 [46,48), layout type = 0 (no swizzle) [61,64)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);
}
// instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): D fp32, A / B bf16, both K-major, dense
__host__ __device__ constexpr uint32_t instr_desc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// the same with the B operand MN-major (bit 16): B is stored [K][N] with N contiguous - core matrices of 8 k-rows x 16 bytes (8
// consecutive n), element (n, k) at (n / 8) * SBO + (k / 8) * LBO + (k % 8) * 16 + (n % 8) * 2 (cute: ((1,n),(8,k)):((X,SBO),(1,LBO)) in
// 16-byte units). This is the layout an epilogue thread that owns ONE accumulator row (= one k of the next product) writes with plain
// 16-byte stores: 8 consecutive columns per store, the 32 lanes of a warp 512 contiguous bytes.
__host__ __device__ constexpr uint32_t instr_desc_bf16_bmn(int m, int n) { return instr_desc_bf16(m, n) | (1u << 16); }

// byte offset of element (r, k) inside an operand tile (see above); k8 = k / 8 (the 16-byte chunk), kc = chunks per row
__device__ __forceinline__ uint32_t chunk_off(int r, int k8, int kc) { return (uint32_t)((r >> 3) * kc * 128 + k8 * 128 + (r & 7) * 16); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t cols) {   // one full warp; cols = power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {          // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// returns false if the phase did not complete within ~`spins` polls (a hang guard: the caller records the failure and goes on)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, uint32_t spins = 1u << 24) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t i = 0; i < spins; i++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// TMA bulk copy global -> shared, completion (bytes) on an mbarrier; `mbar_expect` announces them first. Issued by ONE thread.
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, one K = 16 step; issued by ONE thread
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive fp32 columns of this thread's accumulator row (lane quadrant of the warp) -> registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
// TMEM address of (lane, column) relative to an allocation base: lane in bits [16,32), column in bits [0,16)
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int lane, int col) { return base + ((uint32_t)lane << 16) + (uint32_t)col; }

}   // namespace umma
