// libsplendor_b200.so - fused leaf evaluator: the whole SplendorNNet inference pass in ONE launch.
//
// Replaces, for the tree arena's leaf rows, GenericNNetWrapper.predict (GenericNNetWrapper.py:141-168) +
// SplendorNNet.forward (SplendorNNet.py:127-159): int8 states + legal masks in, exp(log_softmax(masked pi)) and tanh(v)
// out. The torch path (nnet.py) needs ~60 small kernels per wave; here one CTA carries 32 leaves (two groups of 16) through
// every layer with the activations resident in shared memory:
//
//   stage A (the "2d" layers, one row per (leaf, gem column): 112 rows per group, padded to the 128-row UMMA tile)
//       x[56|71|88] -> Linear+BN(7)+ReLU -> Linear+ReLU -> DenseAndPartialGPool(4x8) -> Linear+ReLU
//       on the 5th-generation tensor cores: tcgen05.mma (M = 128, N = 128, bf16 x bf16 -> fp32) issued by one thread, both
//       operands from shared memory in the canonical K-major core-matrix layout (spl_umma.cuh), accumulators in TMEM (one
//       128-column block per group), completion through tcgen05.commit -> mbarrier, epilogues (bias / BatchNorm / ReLU /
//       pooling) by all 16 warps straight from TMEM (tcgen05.ld, one accumulator row per thread) back into the next layer's
//       operand tile. Whole-layer weight blocks stream from L2 into two 36 KB slots by cp.async, two layers ahead.
//   FlattenAndPartialGPool(64, 5)  -> 704 features per leaf
//   stage B (the "1d" layers, 16 rows per group, the 8 warps of a group split the output columns)
//       704 -> 128 -> pool-dense -> 128 -> 128 -> pool-dense -> {PI: 128 -> 406 masked softmax, V: 128 -> n tanh}
//       as warp-level mma.sync m16n8k16 (fragments by ldmatrix, interleaved accumulation chains): 16-row tiles would waste
//       7/8 of a UMMA tile; weights in [n][k] blocks through a four-slot cp.async ring (the same shared memory).
//
// Biases / BatchNorm terms are applied in fp32 in the epilogues. BatchNorm is folded on the host in double precision (eval
// mode); the score-difference head is not evaluated (MCTS never reads it).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "spl_internal.h"
#include "spl_umma.cuh"

namespace {

constexpr int NN_SB = 16;        // leaves per warp group (8 warps)
constexpr int NN_GROUPS = 2;     // warp groups per CTA: both consume the same weight block from shared memory, so a CTA
                                 // streams the network once for 32 leaves (the first version, one group per CTA, pulled
                                 // 190 MB of weights through L2 per 4096-leaf wave)
constexpr int NN_THREADS = 256 * NN_GROUPS;
constexpr int ASTR = 136;        // activation row stride in bf16 elements (128 + 8: conflict-free fragment loads)
constexpr int FLAT_KC = 88;      // 704 flattened features = 88 core matrices per row group of the L4 operand
constexpr int LSTR = 416;        // logits row stride (fp32)
constexpr int NN_ACTIONS = 406;
constexpr int NN_MAXBLK = 48;

// fp32 parameter region of the blob (offsets in floats)
enum {
    P_B1 = 0, P_S1 = 128, P_T1 = 136, P_B2 = 144, P_BG1 = 272, P_SG1 = 400, P_TG1 = 408, P_B3 = 416, P_B4 = 544,
    P_BG4 = 672, P_B5A = 800, P_B5B = 928, P_BG5 = 1056, P_BP0 = 1184, P_BP1 = 1312, P_BV0 = 1760, P_BV1 = 1888,
    P_TOTAL = 1896
};

struct NnPlan {   // byte offsets of the weight blocks inside the blob (after the fp32 parameters)
    int a_off[4], a_bytes[4];   // stage A: whole layers [128 n][K] in the canonical UMMA layout (L1, L2, G1, L3)
    int nblocks;                // stage B: [n][k] blocks with padded rows
    int off[NN_MAXBLK];
    int bytes[NN_MAXBLK];
    int kb[NN_MAXBLK];   // k extent of the block (row stride = kb + 8 elements)
    int total_bytes;
};

inline int kpad1(int n) { return (32 + 10 * n + n * n + 15) / 16 * 16; }

NnPlan make_plan(int n) {
    NnPlan p;
    memset(&p, 0, sizeof p);
    int o = P_TOTAL * 4, b = 0;
    auto add = [&](int nb, int kb) {
        p.off[b] = o; p.kb[b] = kb; p.bytes[b] = nb * (kb + 8) * 2; o += p.bytes[b]; b++;
    };
    const int ka[4] = {kpad1(n), 128, 96, 128};   // L1 dense2d_1.0, L2 dense2d_1.3, G1 partialgpool_1.dense_part.0, L3 dense2d_3.0
    for (int i = 0; i < 4; i++) { p.a_off[i] = o; p.a_bytes[i] = 128 * ka[i] * 2; o += p.a_bytes[i]; }
    for (int i = 0; i < 11; i++) { p.off[b] = o; p.kb[b] = 64; p.bytes[b] = 128 * 64 * 2; o += p.bytes[b]; b++; }   // L4 dense1d_4.0 (704 = 11 x 64): canonical UMMA tiles
    add(64, 112); add(64, 112);                 // G4
    add(64, 128); add(64, 128);                 // L5a
    add(64, 128); add(64, 128);                 // L5b
    add(64, 112); add(64, 112);                 // G5
    add(64, 128); add(64, 128);                 // PI0
    add(64, 128); add(64, 128);                 // V0
    for (int i = 0; i < 7; i++) add(64, 128);   // PI1 (406 -> 448 rows)
    add(64, 128);                               // V1 (n -> 64 rows)
    p.nblocks = b;
    p.total_bytes = o;
    return p;
}

constexpr int SLOT_BYTES = 128 * 72 * 2;   // largest block: [128 n][64 k]
constexpr int L4_SLOTS = 8;                // L4's own ring: the four slots below + the two dead stage A operand tiles (two tiles each)
constexpr int NN_SLOTS = 4;                // weight ring: blocks are requested 3 steps ahead (a step is shorter than an L2 round trip)
static_assert(SLOT_BYTES >= 64 * ASTR * 2, "slot holds a [64][128] block");

constexpr int ATILE_BYTES = 128 * 128 * 2;   // stage A operand tile of a group: 128 rows x 128 k, canonical UMMA layout
struct NnGroupSmem {
    // stage A: the UMMA operand tile (canonical layout); its last layer leaves the activations here as rows of ASTR elements
    // (112 x 136 x 2 = 30,464 bytes) for the flatten step; later the fp32 logits
    __align__(128) __nv_bfloat16 act[ATILE_BYTES / 2];
    __nv_bfloat16 vec[3][NN_SB * ASTR];
};
struct NnSmem {
    // L4 (704 -> 128) runs as ONE M = 128 UMMA tile whose rows 0..31 are the CTA's 32 leaves: the tensor core also reads the 96
    // rows "behind" them (16 row groups x 11,264 bytes = 176 KB from the start of `flat`) and fills accumulator rows nobody
    // looks at. `flat` therefore comes first, so that those reads stay inside this allocation.
    __align__(128) unsigned char flat[4 * FLAT_KC * 128];      // 32 leaves x 704 features, canonical UMMA layout (KC = 88)
    __align__(128) unsigned char slot[NN_SLOTS][SLOT_BYTES];   // stage B: four ring slots; stage A: two slots of 2 SLOT_BYTES (whole layers)
    NnGroupSmem grp[NN_GROUPS];
    float prm[P_TOTAL];                      // biases / BatchNorm terms: read in every epilogue, so not from L2
    uint64_t mma_bar[NN_GROUPS];             // stage A: tcgen05.commit arrives here: one barrier per group's accumulator block
    uint64_t l4_full[L4_SLOTS], l4_empty[L4_SLOTS], l4_done;   // L4: TMA producer -> MMA issuer -> slot free / accumulator ready
    uint32_t tmem_base;
};
static_assert(ATILE_BYTES >= (int)sizeof(__nv_bfloat16) * 7 * NN_SB * ASTR && ATILE_BYTES >= (int)sizeof(float) * NN_SB * LSTR, "aliases of the operand tile");
static_assert(2 * SLOT_BYTES >= 128 * 128 * 2, "a whole 128 x 128 layer fits two ring slots");
static_assert(sizeof(NnSmem) <= 227 * 1024, "shared memory budget");
static_assert(16 * FLAT_KC * 128 <= (int)sizeof(NnSmem), "the rows the tensor core reads behind the 32 real ones lie inside the allocation");

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void sts_bf16x2(__nv_bfloat16* p, float x, float y) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(x, y);
}

// ldmatrix: four 8x8 b16 matrices; lane l supplies the address of row (l & 7) of matrix (l >> 3) and receives, from matrix
// i, the two elements (row l / 4, columns 2 (l % 4), +1) in r[i]
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm2(uint32_t& r0, uint32_t& r1, const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
// lane address of an A fragment (16 rows x 16 k, row-major): matrices = (rows 0-7 | 8-15) x (k 0-7 | 8-15) -> a0..a3
__device__ __forceinline__ const __nv_bfloat16* a_lane_ptr(const __nv_bfloat16* a, int stride, int lane) {
    return a + ((lane & 7) + ((lane >> 3) & 1) * 8) * stride + (lane >> 4) * 8;
}
// lane address of the B fragments of TWO adjacent output tiles (16 n x 16 k of an [n][k] block): matrices = b0, b1 of the
// first tile, b0, b1 of the second
__device__ __forceinline__ const __nv_bfloat16* b2_lane_ptr(const __nv_bfloat16* w, int stride, int lane) {
    return w + ((lane & 7) + (lane >> 4) * 8) * stride + ((lane >> 3) & 1) * 8;
}
// lane address of the B fragments of ONE output tile over TWO k-steps (8 n x 32 k): matrices = b0, b1 of k-step 0, b0, b1 of k-step 1
__device__ __forceinline__ const __nv_bfloat16* b1_lane_ptr(const __nv_bfloat16* w, int stride, int lane) {
    return w + (lane & 7) * stride + (lane >> 3) * 8;
}

// stage B: 16 rows shared by all warps (A from shared memory); this warp computes NT output tiles (8 n each, block rows at
// w[t]) that share the A fragments. mma.sync has a long dependent latency, so the k-steps of every tile alternate between two
// accumulators (summed at the end): 2 NT independent chains of KS / 2 instead of NT chains of KS one after the other.
template <int KS, int NT>
__device__ __forceinline__ void tiles_gemm(float (&c)[NT][4], const __nv_bfloat16* a, int astride, const __nv_bfloat16* const (&w)[NT], int wstride,
                                           int lane) {
    float d[NT][4];
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
        for (int i = 0; i < 4; i++) d[t][i] = 0.f;
    const __nv_bfloat16* al = a_lane_ptr(a, astride, lane);
#pragma unroll
    for (int ks = 0; ks + 1 < KS; ks += 2) {
        uint32_t a0[4], a1[4];
        ldsm4(a0, al + ks * 16);
        ldsm4(a1, al + ks * 16 + 16);
#pragma unroll
        for (int t = 0; t < NT; t++) {
            uint32_t b[4];
            ldsm4(b, b1_lane_ptr(w[t], wstride, lane) + ks * 16);
            mma_bf16(c[t], a0, b[0], b[1]);
            mma_bf16(d[t], a1, b[2], b[3]);
        }
    }
    if (KS & 1) {
        uint32_t a0[4];
        ldsm4(a0, al + (KS - 1) * 16);
#pragma unroll
        for (int t = 0; t < NT; t++) {
            uint32_t b0, b1;
            ldsm2(b0, b1, w[t] + (lane & 7) * wstride + ((lane >> 3) & 1) * 8 + (KS - 1) * 16);
            mma_bf16(c[t], a0, b0, b1);
        }
    }
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
        for (int i = 0; i < 4; i++) c[t][i] += d[t][i];
}
// same A, TWO adjacent output tiles (16 n) of the block; even / odd k-steps go to separate accumulator pairs
template <int KS>
__device__ __forceinline__ void tile2_gemm(float (&c0)[4], float (&c1)[4], float (&d0)[4], float (&d1)[4], const __nv_bfloat16* a, int astride,
                                           const __nv_bfloat16* w, int wstride, int lane) {
    const __nv_bfloat16* al = a_lane_ptr(a, astride, lane);
    const __nv_bfloat16* wl = b2_lane_ptr(w, wstride, lane);
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
        uint32_t af[4], b[4];
        ldsm4(af, al + ks * 16);
        ldsm4(b, wl + ks * 16);
        if (ks & 1) { mma_bf16(d0, af, b[0], b[1]); mma_bf16(d1, af, b[2], b[3]); }
        else { mma_bf16(c0, af, b[0], b[1]); mma_bf16(c1, af, b[2], b[3]); }
    }
}

__device__ long long g_nn_cta[2 * 160];  // diagnostics: globaltimer at the start / end of the first 160 CTAs (slots 32.. of spl_nnet_debug_stamps)
__device__ long long g_nn_stamps[32];   // diagnostics: phase time stamps of CTA 0 (SM clock), read by spl_nnet_debug_stamps
#define NN_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_nn_stamps[i] = clock64(); } while (0)

template <int NP>
__global__ void __launch_bounds__(NN_THREADS, 1) nnet_forward_kernel(const unsigned char* __restrict__ blob, NnPlan plan, const int8_t* __restrict__ states,
                                                                     const uint8_t* __restrict__ valids, const uint8_t* __restrict__ row_src,
                                                                     const int8_t* __restrict__ alt_states, int alt_stride,
                                                                     const uint32_t* __restrict__ alt_mask, int alt_mask_stride, int n_rows,
                                                                     float* __restrict__ pi, float* __restrict__ vout) {
    constexpr int R = 32 + 10 * NP + NP * NP, S = 7 * R, K1 = (R + 15) / 16 * 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    NnSmem& smem_all = *reinterpret_cast<NnSmem*>(smem_raw);
    const float* prm = smem_all.prm;
    for (int i = threadIdx.x; i < P_TOTAL; i += NN_THREADS) smem_all.prm[i] = reinterpret_cast<const float*>(blob)[i];
    const int tid = threadIdx.x, group = tid >> 8, gtid = tid & 255, warp = gtid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    NnGroupSmem& sm = smem_all.grp[group];
    const int base = (blockIdx.x * NN_GROUPS + group) * NN_SB;
    const int live = max(0, min(NN_SB, n_rows - base));
    NN_STAMP(0);
    if (threadIdx.x == 0 && blockIdx.x < 160) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_nn_cta[2 * blockIdx.x] = (long long)gt;
    }

    // ---- stage A weights: whole layers into the two big slots (layer l -> slot l & 1), one cp.async group per layer
    auto issue_a = [&](int l) {
        const unsigned char* src = blob + plan.a_off[l];
        unsigned char* dst = smem_all.slot[2 * (l & 1)];
        for (int i = tid * 16; i < plan.a_bytes[l]; i += NN_THREADS * 16) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    issue_a(0);
    issue_a(1);
    if (tid < 32) umma::tmem_alloc(&smem_all.tmem_base, 256);      // one 128-column accumulator block per group
    if (tid == 32) {
        umma::mbar_init(&smem_all.mma_bar[0], 1); umma::mbar_init(&smem_all.mma_bar[1], 1); umma::mbar_init(&smem_all.l4_done, 1);
        for (int i = 0; i < L4_SLOTS; i++) { umma::mbar_init(&smem_all.l4_full[i], 1); umma::mbar_init(&smem_all.l4_empty[i], 1); }
    }

    // ---- stage B weight-block ring: block b lives in slot b % NN_SLOTS and is requested NN_SLOTS - 1 steps before its use
    auto issue = [&](int b) {
        const unsigned char* src = blob + plan.off[b];
        unsigned char* dst = smem_all.slot[b % NN_SLOTS];
        for (int i = tid * 16; i < plan.bytes[b]; i += NN_THREADS * 16) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    int next_issue = 0;
    // blocks [b, b + count) ready for every thread; the ring stays full (everything before b has been released)
    auto acquire = [&](int b, int count) -> const __nv_bfloat16* {
        const int upto = min(plan.nblocks, b + NN_SLOTS);
        while (next_issue < upto) issue(next_issue++);
        const int pending = next_issue - (b + count);      // groups that may still be in flight
        if (pending >= 3) cp_async_wait<3>();
        else if (pending == 2) cp_async_wait<2>();
        else if (pending == 1) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncthreads();
        return reinterpret_cast<const __nv_bfloat16*>(smem_all.slot[b % NN_SLOTS]);
    };
    auto slot_of = [&](int b) -> const __nv_bfloat16* { return reinterpret_cast<const __nv_bfloat16*>(smem_all.slot[b % NN_SLOTS]); };
    auto release = [&]() { __syncthreads(); };

    // programmatic dependent launch: everything above (parameters, TMEM, barriers, the first weight tiles) may run while the
    // kernel that produces the input rows is still draining; the rows themselves are read only after it has completed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next wave's descent may move in behind this grid
    // the legality bits of the two rows this warp finishes in the softmax (bit j of vbits[q] = action lane + 32 j of row 2 warp + q):
    // requested now, consumed ~30 us later, so the softmax does not start with a global-memory round trip
    uint32_t vbits[2] = {0u, 0u};
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int s = warp * 2 + q;
        if (s < live) {
            if (row_src && row_src[base + s]) {
#pragma unroll
                for (int j = 0; j < 13; j++) vbits[q] |= ((alt_mask[(size_t)j * alt_mask_stride + base + s] >> lane) & 1u) << j;
            } else {
                const uint8_t* va = valids + (size_t)(base + s) * NN_ACTIONS;
#pragma unroll
                for (int j = 0; j < 13; j++)
                    if (lane + 32 * j < NN_ACTIONS) vbits[q] |= (va[lane + 32 * j] != 0 ? 1u : 0u) << j;
            }
        }
    }
    // ---- input: operand tile row c*16 + s, column k = state[s][k][c] (int8 counts are exact in bf16), zero padding up to K1.
    // One thread per (leaf, 8 consecutive k): 56 consecutive state bytes in, seven 16-byte chunks out.
    unsigned char* atile = reinterpret_cast<unsigned char*>(sm.act);
    for (int i = gtid; i < NN_SB * (K1 / 4); i += 256) {   // one thread per (leaf, 4 consecutive k): 28 state bytes in, seven 8-byte half chunks out
        const int s = i / (K1 / 4), k4 = i - s * (K1 / 4);
        const bool row_in = s < live;
        const bool alt = row_in && row_src && row_src[base + s];   // the tree arena's staging row (written by its rules kernel)
        const int8_t* src = (alt ? alt_states + (size_t)(base + s) * alt_stride : states + (size_t)(base + s) * S) + k4 * 28;
        uint32_t packed[7][2];
#pragma unroll
        for (int c = 0; c < 7; c++) packed[c][0] = packed[c][1] = 0u;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const bool in = row_in && (k4 * 4 + kk) < R;
#pragma unroll
            for (int c = 0; c < 7; c++) {
                const float f = in ? (float)src[kk * 7 + c] : 0.f;
                const uint32_t hb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(f));
                packed[c][kk >> 1] |= hb << (16 * (kk & 1));
            }
        }
#pragma unroll
        for (int c = 0; c < 7; c++)
            *reinterpret_cast<uint2*>(atile + umma::chunk_off(c * 16 + s, k4 >> 1, 16) + 8 * (k4 & 1)) = make_uint2(packed[c][0], packed[c][1]);
    }
    NN_STAMP(1);

    // ---- stage A on tcgen05: four layers, per layer  weights landed -> one thread issues the MMAs of both groups -> commit ->
    // everyone waits on the mbarrier -> epilogue from TMEM into the operand tile of the next layer
    {
        const int row = 32 * (warp & 3) + lane;         // accumulator row = TMEM lane of this thread (its warp's lane quadrant)
        const int half = warp >> 2;                     // which 64 of the 128 output columns this thread finishes
        const int c = row >> 4;                         // gem column of the row (BatchNorm1d(7) channel); rows >= 112 are padding
        const bool real = row < 7 * NN_SB;
        const float s1 = real ? prm[P_S1 + c] : 0.f, t1 = real ? prm[P_T1 + c] : 0.f;
        const float sg = real ? prm[P_SG1 + c] : 0.f, tg = real ? prm[P_TG1 + c] : 0.f;
        uint32_t phase = 0;
#pragma unroll 1
        for (int l = 0; l < 4; l++) {
            if (l == 1) NN_STAMP(16);
            if (l == 3) cp_async_wait<0>(); else cp_async_wait<1>();      // this layer's weights (the next layer's may still be in flight)
            umma::fence_smem_to_async();                                  // operand tile (st.shared) and weights -> visible to the tensor core
            umma::fence_before_sync();
            __syncthreads();
            if (l == 1) NN_STAMP(17);
            if (tid == 0) {
                umma::fence_after_sync();
                const int ksteps = l == 0 ? K1 / 16 : (l == 2 ? 6 : 8);
                const uint32_t a_skip = l == 2 ? 4 * 128 : 0;             // G1 reads columns 32..127 of the tile
                // descriptors of k-step 0; step k starts 256 bytes further = +16 in the (address >> 4) field
                const uint64_t bd0 = umma::smem_desc(umma::smem_u32(smem_all.slot[2 * (l & 1)]), 128, 2 * ksteps * 128);
                const uint32_t tb = smem_all.tmem_base;
#pragma unroll
                for (int gi = 0; gi < NN_GROUPS; gi++) {
                    const uint64_t ad0 = umma::smem_desc(umma::smem_u32(smem_all.grp[gi].act) + a_skip, 128, 16 * 128);
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if (k < ksteps) umma::mma_bf16_ss(tb + gi * 128, ad0 + (uint64_t)(16 * k), bd0 + (uint64_t)(16 * k), umma::instr_desc_bf16(128, 128), k > 0);
                    umma::commit(&smem_all.mma_bar[gi]);                  // group 0's epilogue runs under group 1's MMAs
                }
            }
            if (l == 1) NN_STAMP(18);
            umma::mbar_wait(&smem_all.mma_bar[group], phase);
            umma::fence_after_sync();
            if (l == 1) NN_STAMP(19);
            const uint32_t trow = umma::tmem_addr(smem_all.tmem_base, 32 * (warp & 3), group * 128 + half * 64);
#pragma unroll
            for (int part = 0; part < 2; part++) {
                float v[32];
                __syncwarp();                                             // tcgen05.ld is warp-collective: padding rows rejoin here
                umma::tmem_ld32(trow + part * 32, v);
                const int n0 = half * 64 + part * 32;
                if (real) {
                uint32_t pk[16];
                // bias (+ BatchNorm1d(7) of the row's gem column) + ReLU; biases four at a time (the G1 bias block is followed by
                // its BatchNorm terms, which the dropped columns 120..127 read harmlessly)
                const int pb = l == 0 ? P_B1 : (l == 1 ? P_B2 : (l == 2 ? P_BG1 : P_B3));
                const float sc = l == 0 ? s1 : (l == 2 ? sg : 1.f), sh = l == 0 ? t1 : (l == 2 ? tg : 0.f);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(prm + pb + n0 + j);
                    const __nv_bfloat162 h01 = __floats2bfloat162_rn(fmaxf((v[j] + b4.x) * sc + sh, 0.f), fmaxf((v[j + 1] + b4.y) * sc + sh, 0.f));
                    const __nv_bfloat162 h23 = __floats2bfloat162_rn(fmaxf((v[j + 2] + b4.z) * sc + sh, 0.f), fmaxf((v[j + 3] + b4.w) * sc + sh, 0.f));
                    pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h01);
                    pk[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                }
                if (l == 1 && n0 == 0) {
                    // the pooled half of DenseAndPartialGPool: max and mean of the 4 groups of 8 leading columns (of the bf16
                    // activations) replace columns 0..7, where the next layer leaves them next to its 120 dense outputs
                    uint32_t pool[4];
                    float mxs[4], avs[4];
#pragma unroll
                    for (int gq = 0; gq < 4; gq++) {
                        float mx = -INFINITY, sum = 0.f;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk[gq * 4 + q]));
                            mx = fmaxf(mx, fmaxf(f.x, f.y)); sum += f.x + f.y;
                        }
                        mxs[gq] = mx; avs[gq] = sum * 0.125f;
                    }
                    const __nv_bfloat162 m01 = __floats2bfloat162_rn(mxs[0], mxs[1]), m23 = __floats2bfloat162_rn(mxs[2], mxs[3]);
                    const __nv_bfloat162 a01 = __floats2bfloat162_rn(avs[0], avs[1]), a23 = __floats2bfloat162_rn(avs[2], avs[3]);
                    pool[0] = *reinterpret_cast<const uint32_t*>(&m01); pool[1] = *reinterpret_cast<const uint32_t*>(&m23);
                    pool[2] = *reinterpret_cast<const uint32_t*>(&a01); pool[3] = *reinterpret_cast<const uint32_t*>(&a23);
                    pk[0] = pool[0]; pk[1] = pool[1]; pk[2] = pool[2]; pk[3] = pool[3];
                }
                if (l == 3) {           // last layer: rows of ASTR elements for the flatten step
                    uint4* dst = reinterpret_cast<uint4*>(sm.act + row * ASTR + n0);
#pragma unroll
                    for (int q = 0; q < 4; q++) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                } else {
                    const int chunk0 = l == 2 ? 1 + n0 / 8 : n0 / 8;      // G1 writes its outputs to columns 8 + n
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (chunk0 + q < 16)
                            *reinterpret_cast<uint4*>(atile + umma::chunk_off(row, chunk0 + q, 16)) =
                                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
                }
            }
            if (l == 1) NN_STAMP(20);
            if (group == 0) umma::mbar_wait(&smem_all.mma_bar[1], phase);   // every MMA of the layer is done: its weight slot is free again
            phase ^= 1u;
            if (l < 2) issue_a(l + 2);
            if (l == 3 && tid == 0)                                       // L4's first four weight tiles travel from here on
                for (int j = 0; j < NN_SLOTS; j++) {
                    umma::mbar_expect(&smem_all.l4_full[j], 128 * 64 * 2);
                    umma::bulk_load(smem_all.slot[j], blob + plan.off[j], 128 * 64 * 2, &smem_all.l4_full[j]);
                }
            NN_STAMP(2 + l);
        }
        umma::fence_before_sync();
        __syncthreads();
    }

    // ---- FlattenAndPartialGPool(64, 5): [max over the 5 gem colours | mean | gold, points rows | last 64 features of all 7]
    // -> row (group * 16 + leaf) of the L4 operand (canonical UMMA layout, 88 core matrices per row group)
    auto flat_at = [&](int r, int k) -> __nv_bfloat16* {
        return reinterpret_cast<__nv_bfloat16*>(smem_all.flat + umma::chunk_off(r, k >> 3, FLAT_KC) + (k & 7) * 2);
    };
    for (int i = gtid; i < NN_SB * 64; i += 256) {          // first 64 features of every row: pooled over the colour rows
        const int s = i >> 6, j = i & 63, r = group * NN_SB + s;
        float a[7];
#pragma unroll
        for (int c = 0; c < 7; c++) a[c] = __bfloat162float(sm.act[(c * 16 + s) * ASTR + j]);
        const float mx = fmaxf(fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3])), a[4]);
        const float sum = a[0] + a[1] + a[2] + a[3] + a[4];
        *flat_at(r, j) = __float2bfloat16(mx); *flat_at(r, 64 + j) = __float2bfloat16(sum * 0.2f);
        *flat_at(r, 128 + j) = __float2bfloat16(a[5]); *flat_at(r, 192 + j) = __float2bfloat16(a[6]);
    }
    for (int i = gtid; i < NN_SB * 7 * 8; i += 256) {       // last 64 features of all 7 rows: copied, one 16-byte chunk (8 features) at a time
        const int s = i / 56, rr = i - s * 56, c = rr >> 3, j8 = rr & 7;
        *reinterpret_cast<uint4*>(flat_at(group * NN_SB + s, 256 + c * 64 + j8 * 8)) = *reinterpret_cast<const uint4*>(sm.act + (c * 16 + s) * ASTR + 64 + j8 * 8);
    }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();

    NN_STAMP(6);
    // ---- L4: Linear(704,128) + ReLU on tcgen05 as a three-role pipeline over the four ring slots: one thread streams the eleven
    // [128 n][64 k] weight tiles by TMA bulk copies (mbarrier complete_tx), one thread issues four MMAs per tile as it lands and
    // frees the slot with tcgen05.commit, everybody else sleeps on the barrier of the finished accumulator (TMEM columns 0..127).
    // Eight slots: the four ring slots (their tiles are requested as soon as stage A's last MMAs are done, i.e. they travel during
    // that layer's epilogue and the flatten step) and the two operand tiles of stage A, which are dead once the flatten step has
    // read them (a 16 KB tile takes ~1 us to arrive, an MMA group 0.13 us: the depth of the ring is what sets the pace).
    {
        const int wc = tid >> 5;
        auto l4_slot = [&](int i) -> unsigned char* {
            return i < NN_SLOTS ? smem_all.slot[i] : reinterpret_cast<unsigned char*>(smem_all.grp[(i - NN_SLOTS) >> 1].act) + ((i - NN_SLOTS) & 1) * (128 * 64 * 2);
        };
        if (wc == 0) {
            if (lane == 0) {
                for (int j = NN_SLOTS; j < 11; j++) {                     // tiles 0..3 were requested after stage A's last MMAs
                    const int sl = j % L4_SLOTS;
                    if (j >= L4_SLOTS) umma::mbar_wait(&smem_all.l4_empty[sl], (uint32_t)((j / L4_SLOTS - 1) & 1));
                    umma::mbar_expect(&smem_all.l4_full[sl], 128 * 64 * 2);
                    umma::bulk_load(l4_slot(sl), blob + plan.off[j], 128 * 64 * 2, &smem_all.l4_full[sl]);
                }
            }
            __syncwarp();
        } else if (wc == 1) {
            if (lane == 0) {
                umma::fence_after_sync();
                const uint32_t idesc = umma::instr_desc_bf16(128, 128);
                const uint32_t tb = smem_all.tmem_base;
                for (int j = 0; j < 11; j++) {
                    const int sl = j % L4_SLOTS;
                    umma::mbar_wait(&smem_all.l4_full[sl], (uint32_t)((j / L4_SLOTS) & 1));
                    umma::fence_after_sync();
                    const uint64_t ad0 = umma::smem_desc(umma::smem_u32(smem_all.flat) + j * 8 * 128, 128, FLAT_KC * 128);
                    const uint64_t bd0 = umma::smem_desc(umma::smem_u32(l4_slot(sl)), 128, 8 * 128);
#pragma unroll
                    for (int k = 0; k < 4; k++) umma::mma_bf16_ss(tb, ad0 + (uint64_t)(16 * k), bd0 + (uint64_t)(16 * k), idesc, j > 0 || k > 0);
                    umma::commit(&smem_all.l4_empty[sl]);
                }
                umma::commit(&smem_all.l4_done);
            }
            __syncwarp();
        }
        umma::mbar_wait(&smem_all.l4_done, 0);
        umma::fence_after_sync();
        if ((wc & 3) == 0) {      // the four warps that own TMEM lanes 0..31: lane = leaf, 32 output columns each -> bias + ReLU -> vec[0]
            float v[32];
            umma::tmem_ld32(umma::tmem_addr(smem_all.tmem_base, 0, 32 * (wc >> 2)), v);
            __nv_bfloat16* o = smem_all.grp[lane >> 4].vec[0] + (lane & 15) * ASTR + 32 * (wc >> 2);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const float4 b0 = *reinterpret_cast<const float4*>(prm + P_B4 + 32 * (wc >> 2) + j);
                const float4 b1 = *reinterpret_cast<const float4*>(prm + P_B4 + 32 * (wc >> 2) + j + 4);
                const __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(v[j] + b0.x, 0.f), fmaxf(v[j + 1] + b0.y, 0.f));
                const __nv_bfloat162 h1 = __floats2bfloat162_rn(fmaxf(v[j + 2] + b0.z, 0.f), fmaxf(v[j + 3] + b0.w, 0.f));
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(v[j + 4] + b1.x, 0.f), fmaxf(v[j + 5] + b1.y, 0.f));
                const __nv_bfloat162 h3 = __floats2bfloat162_rn(fmaxf(v[j + 6] + b1.z, 0.f), fmaxf(v[j + 7] + b1.w, 0.f));
                *reinterpret_cast<uint4*>(o + j) = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                                              *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        if (tid < 32) umma::tmem_dealloc(smem_all.tmem_base, 256);
    }
    int blk = 11;          // the remaining stage B tiles go through the cp.async ring (all four slots are free again)
    next_issue = 11;
    // stage B helpers: in -> out, the two [64 n][KB] blocks of a layer in ONE ring step, warp w owns tile w of each block
    auto dense_vec = [&](const __nv_bfloat16* in, __nv_bfloat16* out, int pbias, bool relu) {
        acquire(blk, 2);
        if (pbias == P_B5A) NN_STAMP(12);
        float cc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        {
            const __nv_bfloat16* const ws[2] = {slot_of(blk) + warp * 8 * ASTR, slot_of(blk + 1) + warp * 8 * ASTR};
            tiles_gemm<8, 2>(cc, in, ASTR, ws, ASTR, lane);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float(&c)[4] = cc[h];
            const int n = h * 64 + warp * 8 + 2 * t;
            const float b0 = prm[pbias + n], b1 = prm[pbias + n + 1];
            float v0 = c[0] + b0, v1 = c[1] + b1, v2 = c[2] + b0, v3 = c[3] + b1;
            if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
            sts_bf16x2(out + g * ASTR + n, v0, v1);
            sts_bf16x2(out + (g + 8) * ASTR + n, v2, v3);
        }
        if (pbias == P_B5A) NN_STAMP(13);
        release(); blk += 2;
        if (pbias == P_B5A) NN_STAMP(14);
    };
    auto pool_dense_vec = [&](const __nv_bfloat16* in, __nv_bfloat16* out, int pbias) {   // 4 groups of 4 + Linear(112,120)+BN(1)+ReLU
        if (gtid < 64) {
            const int row = gtid >> 2, grp = gtid & 3;
            const uint2 raw = *reinterpret_cast<const uint2*>(in + row * ASTR + grp * 4);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
            const float2 a = __bfloat1622float2(h2[0]), b = __bfloat1622float2(h2[1]);
            out[row * ASTR + grp] = __float2bfloat16(fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y)));
            out[row * ASTR + 4 + grp] = __float2bfloat16((a.x + a.y + b.x + b.y) * 0.25f);
        }
        acquire(blk, 2);
        float cc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        {
            const __nv_bfloat16* const ws[2] = {slot_of(blk) + warp * 8 * 120, slot_of(blk + 1) + warp * 8 * 120};
            tiles_gemm<7, 2>(cc, in + 16, ASTR, ws, 120, lane);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float(&c)[4] = cc[h];
            const int n = h * 64 + warp * 8 + 2 * t;
            if (n < 120) {
                const float b0 = prm[pbias + n], b1 = prm[pbias + n + 1];
                sts_bf16x2(out + g * ASTR + 8 + n, fmaxf(c[0] + b0, 0.f), fmaxf(c[1] + b1, 0.f));
                sts_bf16x2(out + (g + 8) * ASTR + 8 + n, fmaxf(c[2] + b0, 0.f), fmaxf(c[3] + b1, 0.f));
            }
        }
        release(); blk += 2;
    };
    NN_STAMP(7);
    pool_dense_vec(sm.vec[0], sm.vec[1], P_BG4);
    dense_vec(sm.vec[1], sm.vec[0], P_B5A, true);
    dense_vec(sm.vec[0], sm.vec[1], P_B5B, true);
    pool_dense_vec(sm.vec[1], sm.vec[0], P_BG5);
    dense_vec(sm.vec[0], sm.vec[1], P_BP0, false);   // output_layers_PI.0 (no activation)
    dense_vec(sm.vec[0], sm.vec[2], P_BV0, false);   // output_layers_V.0

    NN_STAMP(8);
    // ---- PI1: 128 -> 406 logits (fp32, in the former activation buffer), 7 blocks of 64 rows
    float* logits = reinterpret_cast<float*>(sm.act);
    for (int h = 0; h < 7; h += 2) {   // two 64-row blocks per ring step
        const int cnt = h + 1 < 7 ? 2 : 1;
        acquire(blk, cnt);
        float cc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        {   // (the last step has one block: its second tile re-reads the first and is dropped)
            const __nv_bfloat16* const ws[2] = {slot_of(blk) + warp * 8 * ASTR, slot_of(blk + cnt - 1) + warp * 8 * ASTR};
            tiles_gemm<8, 2>(cc, sm.vec[1], ASTR, ws, ASTR, lane);
        }
        for (int q = 0; q < cnt; q++) {
            float(&c)[4] = cc[q];
            const int n = (h + q) * 64 + warp * 8 + 2 * t;
            if (n < LSTR) {
                const float b0 = prm[P_BP1 + n], b1 = prm[P_BP1 + n + 1];
                logits[g * LSTR + n] = c[0] + b0; logits[g * LSTR + n + 1] = c[1] + b1;
                logits[(g + 8) * LSTR + n] = c[2] + b0; logits[(g + 8) * LSTR + n + 1] = c[3] + b1;
            }
        }
        release(); blk += cnt;
    }
    NN_STAMP(9);
    // ---- V1: 128 -> n, tanh
    {
        const __nv_bfloat16* w = acquire(blk, 1);
        if (warp == 0) {
            float cc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
            const __nv_bfloat16* const ws[1] = {w};
            tiles_gemm<8, 1>(cc, sm.vec[2], ASTR, ws, ASTR, lane);
            float(&c)[4] = cc[0];
            const int n = 2 * t;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int row = g + (q >> 1) * 8, col = n + (q & 1);
                if (col < NP && row < live) vout[(size_t)(base + row) * NP + col] = tanhf(c[q] + prm[P_BV1 + col]);
            }
        }
        release(); blk++;
    }
    NN_STAMP(10);
    // ---- masked softmax: log_softmax(where(valid, pi, -1e8)) then exp (SplendorNNet.py:153-159, GenericNNetWrapper.py:166)
    for (int s = warp * 2; s < warp * 2 + 2; s++) {
        if (s >= live) continue;
        const uint32_t vb = vbits[s - warp * 2];
        float x[13], mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 13; j++) {
            const int a = lane + 32 * j;
            x[j] = -INFINITY;
            if (a < NN_ACTIONS) {
                x[j] = ((vb >> j) & 1u) ? logits[s * LSTR + a] : -1e8f;
                mx = fmaxf(mx, x[j]);
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 13; j++) {
            x[j] = (lane + 32 * j) < NN_ACTIONS ? expf(x[j] - mx) : 0.f;
            sum += x[j];
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < 13; j++) {
            const int a = lane + 32 * j;
            if (a < NN_ACTIONS) pi[(size_t)(base + s) * NN_ACTIONS + a] = x[j] * inv;
        }
    }
    NN_STAMP(11);
    if (threadIdx.x == 0 && blockIdx.x < 160) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_nn_cta[2 * blockIdx.x + 1] = (long long)gt;
    }
}

// ------------------------------------------------------------------------------------------ host: BatchNorm folding + blob packing
struct BnFold { double s[8], t[8]; };
BnFold fold_bn(const float* w, const float* b, const float* mean, const float* var, int c) {
    BnFold f;
    for (int i = 0; i < c; i++) {
        f.s[i] = (double)w[i] / sqrt((double)var[i] + 1e-5);
        f.t[i] = (double)b[i] - (double)mean[i] * f.s[i];
    }
    return f;
}
uint16_t to_bf16(double x) {
    float f = (float)x;
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
    u += 0x7FFFu + ((u >> 16) & 1u);   // round to nearest even
    return (uint16_t)(u >> 16);
}
// writes rows [n0, n0+nb) x cols [k0, k0+kb) of W[N][K] (scaled by `scale`) as a padded [nb][kb+8] bf16 block
void pack_block(unsigned char* dst, const float* W, int N, int K, int n0, int nb, int k0, int kb, double scale) {
    uint16_t* d = reinterpret_cast<uint16_t*>(dst);
    for (int r = 0; r < nb; r++)
        for (int c = 0; c < kb + 8; c++) {
            const int n = n0 + r, k = k0 + c;
            d[r * (kb + 8) + c] = (n < N && c < kb && k < K) ? to_bf16((double)W[(size_t)n * K + k] * scale) : (uint16_t)0;
        }
}

// rows [n0, n0 + nb) x columns [k0, k0 + kb) of W[N][K] (zero outside W) as a [nb][kb] bf16 tile in the canonical UMMA K-major
// layout (spl_umma.cuh): element (r, c) at byte (r / 8) * (kb / 8) * 128 + (c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2
void pack_canonical(unsigned char* dst, const float* W, int N, int K, int n0, int nb, int k0, int kb) {
    uint16_t* d = reinterpret_cast<uint16_t*>(dst);
    const int kc = kb / 8;
    for (int r = 0; r < nb; r++)
        for (int c = 0; c < kb; c++) {
            const int n = n0 + r, k = k0 + c;
            const size_t off = ((size_t)(r >> 3) * kc * 128 + (size_t)(c >> 3) * 128 + (r & 7) * 16 + (c & 7) * 2) / 2;
            d[off] = (n < N && k < K) ? to_bf16((double)W[(size_t)n * K + k]) : (uint16_t)0;
        }
}

}   // namespace

int spl_nnet_impl_() {
    static const int impl = [] { const char* e = getenv("SPL_NNET_IMPL"); return (e && e[0] == '1') ? 1 : 2; }();
    return impl;
}

extern "C" {

size_t spl_nnet_blob_bytes(int n_players) {
    if (n_players < 2 || n_players > 4) return 0;
    if (spl_nnet_impl_() == 2) return nn2::blob_bytes(n_players);
    return (size_t)make_plan(n_players).total_bytes;
}

int spl_nnet_pack(int n_players, const float* const* T, void* blob, size_t blob_bytes) {
    if (n_players < 2 || n_players > 4 || !T || !blob) return spl_fail_(SPL_E_ARG, "spl_nnet_pack: bad argument");
    if (spl_nnet_impl_() == 2) {
        for (int i = 0; i < 46; i++)
            if (!T[i]) return spl_fail_(SPL_E_ARG, "spl_nnet_pack: null tensor");
        return nn2::pack(n_players, T, blob, blob_bytes);
    }
    const NnPlan p = make_plan(n_players);
    if (blob_bytes < (size_t)p.total_bytes) return spl_fail_(SPL_E_ARG, "spl_nnet_pack: blob smaller than spl_nnet_blob_bytes");
    for (int i = 0; i < 46; i++)
        if (!T[i]) return spl_fail_(SPL_E_ARG, "spl_nnet_pack: null tensor");
    const int R = 32 + 10 * n_players + n_players * n_players, K1 = kpad1(n_players);
    unsigned char* B = (unsigned char*)blob;
    memset(B, 0, p.total_bytes);
    float* prm = reinterpret_cast<float*>(B);
    const BnFold bn1 = fold_bn(T[2], T[3], T[4], T[5], 7), bng1 = fold_bn(T[10], T[11], T[12], T[13], 7);
    const BnFold bn4 = fold_bn(T[20], T[21], T[22], T[23], 1), bn5 = fold_bn(T[26], T[27], T[28], T[29], 1), bng5 = fold_bn(T[34], T[35], T[36], T[37], 1);
    for (int i = 0; i < 128; i++) {
        prm[P_B1 + i] = T[1][i]; prm[P_B2 + i] = T[7][i]; prm[P_B3 + i] = T[15][i]; prm[P_B4 + i] = T[17][i];
        prm[P_B5A + i] = (float)((double)T[25][i] * bn5.s[0] + bn5.t[0]);
        prm[P_B5B + i] = T[31][i]; prm[P_BP0 + i] = T[39][i]; prm[P_BV0 + i] = T[43][i];
    }
    for (int i = 0; i < 120; i++) {
        prm[P_BG1 + i] = T[9][i];
        prm[P_BG4 + i] = (float)((double)T[19][i] * bn4.s[0] + bn4.t[0]);
        prm[P_BG5 + i] = (float)((double)T[33][i] * bng5.s[0] + bng5.t[0]);
    }
    for (int i = 0; i < 7; i++) {
        prm[P_S1 + i] = (float)bn1.s[i]; prm[P_T1 + i] = (float)bn1.t[i];
        prm[P_SG1 + i] = (float)bng1.s[i]; prm[P_TG1 + i] = (float)bng1.t[i];
    }
    for (int i = 0; i < NN_ACTIONS; i++) prm[P_BP1 + i] = T[41][i];
    for (int i = 0; i < n_players; i++) prm[P_BV1 + i] = T[45][i];
    int b = 0;
    auto two = [&](const float* W, int N, int K, int kb, double scale) {   // two [64][kb] blocks: n halves
        pack_block(B + p.off[b], W, N, K, 0, 64, 0, kb, scale); b++;
        pack_block(B + p.off[b], W, N, K, 64, 64, 0, kb, scale); b++;
    };
    pack_canonical(B + p.a_off[0], T[0], 128, R, 0, 128, 0, K1);
    pack_canonical(B + p.a_off[1], T[6], 128, 128, 0, 128, 0, 128);
    pack_canonical(B + p.a_off[2], T[8], 120, 96, 0, 128, 0, 96);
    pack_canonical(B + p.a_off[3], T[14], 128, 128, 0, 128, 0, 128);
    for (int kbk = 0; kbk < 11; kbk++) { pack_canonical(B + p.off[b], T[16], 128, 704, 0, 128, kbk * 64, 64); b++; }
    two(T[18], 120, 112, 112, bn4.s[0]);
    two(T[24], 128, 128, 128, bn5.s[0]);
    two(T[30], 128, 128, 128, 1.0);
    two(T[32], 120, 112, 112, bng5.s[0]);
    two(T[38], 128, 128, 128, 1.0);
    two(T[42], 128, 128, 128, 1.0);
    for (int h = 0; h < 7; h++) { pack_block(B + p.off[b], T[40], NN_ACTIONS, 128, h * 64, 64, 0, 128, 1.0); b++; }
    pack_block(B + p.off[b], T[44], n_players, 128, 0, 64, 0, 128, 1.0); b++;
    if (b != p.nblocks) return spl_fail_(SPL_E_ARG, "spl_nnet_pack: internal block count mismatch");
    return SPL_OK;
}

int spl_nnet_debug_stamps(long long* out32) {   /* diagnostics only: SM-clock stamps of the last launch's CTA 0 */
    if (spl_nnet_impl_() == 2) return nn2::debug_stamps(out32);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out32, g_nn_stamps, sizeof(long long) * 32));
    return SPL_OK;
}

int spl_nnet_debug_tile_stamps(long long* out144) {   /* diagnostics only: per weight tile of CTA 0: requested / landed / MMAs issued (SM clock) */
    if (spl_nnet_impl_() != 2) return spl_fail_(SPL_E_ARG, "spl_nnet_debug_tile_stamps: transposed evaluator only");
    return nn2::debug_tile_stamps(out144);
}

int spl_nnet_debug_cta_times(long long* out320) {   /* diagnostics only: globaltimer (ns) at the start / end of the first 160 CTAs */
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out320, g_nn_cta, sizeof(long long) * 320));
    return SPL_OK;
}

int spl_nnet_forward(spl_ctx* c, const void* blob, const int8_t* states, const uint8_t* valids, int n_rows, float* pi, float* v, void* stream) {
    return spl_nnet_forward_rows_(c, blob, states, valids, nullptr, nullptr, 0, nullptr, 0, n_rows, pi, v, (cudaStream_t)stream, false);
}

}   // extern "C"

int spl_nnet_forward_rows_(spl_ctx* c, const void* blob, const int8_t* states, const uint8_t* valids, const uint8_t* row_src,
                           const int8_t* alt_states, int alt_stride, const uint32_t* alt_mask, int alt_mask_stride, int n_rows, float* pi,
                           float* v, cudaStream_t st, bool programmatic_dependent) {
    if (!c) return spl_fail_(SPL_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (!blob || !states || !valids || !pi || !v || n_rows <= 0) return spl_fail_(SPL_E_ARG, "spl_nnet_forward: bad argument");
    if (row_src && (!alt_states || !alt_mask)) return spl_fail_(SPL_E_ARG, "spl_nnet_forward: row_src without staging rows");
    if (((uintptr_t)blob & 15u) != 0) return spl_fail_(SPL_E_ARG, "spl_nnet_forward: blob must be 16-byte aligned");
    if (spl_nnet_impl_() == 2)
        return nn2::forward_rows(c, blob, states, valids, row_src, alt_states, alt_stride, alt_mask, alt_mask_stride, n_rows, pi, v, st, programmatic_dependent);
    const NnPlan p = make_plan(c->n);
    const int grid = (n_rows + NN_SB * NN_GROUPS - 1) / (NN_SB * NN_GROUPS);
    const int smem = (int)sizeof(NnSmem);
    DISPATCH_N(c->n, {
        auto k = nnet_forward_kernel<N>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NN_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cfg.attrs = attr; cfg.numAttrs = programmatic_dependent ? 1 : 0;
        CU(cudaLaunchKernelEx(&cfg, k, (const unsigned char*)blob, p, states, valids, row_src, alt_states, alt_stride, alt_mask, alt_mask_stride, n_rows, pi, v));
    });
    CU(cudaGetLastError());
    return SPL_OK;
}
