// libsplendor_b200.so - fused leaf evaluator, transposed form: the whole SplendorNNet inference pass in ONE launch, every layer
// on the 5th-generation tensor cores.
//
// Replaces, for the tree arena's leaf rows, GenericNNetWrapper.predict (GenericNNetWrapper.py:141-168) +
// SplendorNNet.forward (SplendorNNet.py:127-159): int8 states + legal masks in, exp(log_softmax(masked pi)) and tanh(v) out.
//
// Every layer is computed as  D^T[out feature][column] = W[out feature][k] . X^T[k][column]:
//   * the WEIGHTS are the A operand (M = 128 output features = the 128 TMEM lanes), streamed from L2 as [128][<= 64 k] tiles in the
//     canonical K-major layout by one producer thread (cp.async.bulk + mbarrier complete_tx) through a six-slot ring;
//   * the ACTIVATIONS are the B operand, held MN-major in shared memory ([k][column], 8 consecutive columns = 16 bytes), N = the
//     columns themselves: 7 x 32 = 224 (leaf, gem column) pairs per half in the "2d" layers, 64 leaves in the "1d" layers. No row
//     of any MMA is padding (the first version put the activations in the M dimension: 112 of 128 rows in the 2d layers, 32 of
//     128 in the 704-wide layer, and ran the small layers on mma.sync).
//   * an epilogue thread owns ONE accumulator row (TMEM lane = output feature) and a range of columns: bias is one register,
//     BatchNorm1d(7) depends on the column index (compile time), and its 16-byte stores of 8 consecutive columns land exactly
//     in the next layer's B operand (the 32 lanes of a warp write 512 contiguous bytes: no bank conflicts, no transposition).
// A CTA carries 64 leaves: in the 2d layers as two halves of 32 whose MMAs and epilogues alternate (the tensor pipe works on one
// half while all 16 epilogue warps work on the other), in the 1d layers as one N = 64 chain. Roles: warps 0..15 epilogues,
// warps 16, 17 weight producers, warps 18 / 19 MMA issuers (A: the 2d layers; A and B share the 1d layers); they meet only on mbarriers.
//
// The pooled halves of DenseAndPartialGPool (SplendorNNet.py:6-29; max / mean over groups of OUTPUT FEATURES, i.e. across TMEM
// lanes) are computed from the shared-memory operand by a few threads into a 16-row side operand and enter the next layer as
// one extra K = 16 step. FlattenAndPartialGPool (SplendorNNet.py:32-54) happens in the registers of the last 2d epilogue (a
// thread holds its feature for all 7 gem columns of a leaf) and is written straight into the 704-row operand of the next layer.
// output_layers_PI / _V are two Linear layers without an activation in between (SplendorNNet.py:101-109): folded into one
// [406 + n][128] matrix in double precision on the host; the value rows ride in the last M tile of the policy head.
// BatchNorm is folded on the host in double precision (eval mode); the score-difference head is not evaluated (MCTS never reads it).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "spl_internal.h"
#include "spl_umma.cuh"

namespace nn2 {

constexpr int NL = 64;                 // leaves per CTA
constexpr int HL = 32;                 // leaves per half (2d layers)
constexpr int HC = 7 * HL;             // columns per half: (leaf, gem column) pairs, column = 7 * leaf + gem column
constexpr int RING = 6;                // weight tile slots
constexpr int TILE = 128 * 64 * 2;     // a full weight tile: [128 m][64 k] bf16
constexpr int KSTEP_BYTES = 128 * 16 * 2;
constexpr int OPER_BYTES = 2 * (HC / 8) * 2048;   // 2d operand [128 k][448 columns]: element (k, n) at (n / 8) * 2048 + k * 16 + (n % 8) * 2
constexpr int HALF_BYTES = (HC / 8) * 2048;
constexpr int FLAT_SBO = 88 * 128;     // 704-row operand [704 k][64 leaves]: (leaf / 8) * 11264 + k * 16 + (leaf % 8) * 2 (aliases the 2d operand)
constexpr int VEC_OFF = 8 * FLAT_SBO;  // 1d operand [128 k][64 leaves] behind it: (leaf / 8) * 2048 + k * 16 + (leaf % 8) * 2
constexpr int POOL_BYTES = 2 * (HC / 8) * 256;    // side operand [16 k][448 columns]: (n / 8) * 256 + k * 16 + (n % 8) * 2; rows 8..15 stay zero
constexpr int LSTR = 416;              // fp32 logits row stride (aliases the operand region once the head's MMAs are done)
constexpr int NN_ACTIONS = 406;
constexpr int EPI_THREADS = 512;       // warps 0..15: epilogues
constexpr int PRODUCERS = 2;           // warps 16, 17: weight producers (bulk copies of ONE thread complete one after the other, ~840 cycles
                                       // each whatever their size - profiles/tools/stream_cycles.py - so tile t is requested by warp 16 + t % 2;
                                       // 20 warps in all: a 21st would cut the register budget from 96 to 80 per thread - five warps per
                                       // sub-partition instead of six)
constexpr int MMA_WARP = 16 + PRODUCERS;   // issuer A: the 2d layers and its share of the 1d layers
constexpr int MMA_WARP_B = MMA_WARP + 1;   // issuer B: the other share of the 1d layers (the cost of a weight tile is the issuing thread's
                                           // time - wait ~180 cycles, descriptors, 4 x 86 cycles of issue - not the tensor pipe's)
constexpr int THREADS = 32 * (MMA_WARP_B + 1);
constexpr int MAX_TILES = 48;
static_assert(VEC_OFF + 8 * 2048 <= OPER_BYTES, "1d operand inside the region");
static_assert(NL * LSTR * 4 <= OPER_BYTES, "logits inside the region");

// fp32 parameter region at the start of the blob (offsets in floats)
enum { P_B1 = 0, P_B2 = 128, P_BG1 = 256, P_B3 = 384, P_B4 = 512, P_BG4 = 640, P_B5A = 768, P_B5B = 896, P_BG5 = 1024, P_BH = 1152,
       P_S1 = 1664, P_T1 = 1672, P_SG1 = 1680, P_TG1 = 1688, P_TOTAL = 1696 };

struct Plan {
    int nt, nt_l1;
    int off[MAX_TILES], bytes[MAX_TILES];
    int total_bytes;
};

inline int kpad1(int n) { return (32 + 10 * n + n * n + 15) / 16 * 16; }

Plan make_plan(int n) {
    Plan p;
    memset(&p, 0, sizeof p);
    int o = P_TOTAL * 4, t = 0;
    auto add = [&](int ksteps) { p.off[t] = o; p.bytes[t] = ksteps * KSTEP_BYTES; o += p.bytes[t]; t++; };
    const int k1 = kpad1(n) / 16;
    add(k1 < 4 ? k1 : 4);
    if (k1 > 4) add(k1 - 4);
    p.nt_l1 = t;
    add(4); add(4);                                   // L2   dense2d_1.3
    add(4); add(2);                                   // G1   partialgpool_1.dense_part.0 (features 32..127)
    add(4); add(4); add(1);                           // L3   dense2d_3.0: 120 dense inputs + the 8 pooled ones
    for (int i = 0; i < 11; i++) add(4);              // L4   dense1d_4.0 (704 = 11 x 64)
    add(4); add(3);                                   // G4   partialgpool_4.dense_part.0 (features 16..127)
    add(4); add(4); add(1);                           // L5a  dense1d_5.0
    add(4); add(4);                                   // L5b  dense1d_5.3
    add(4); add(3);                                   // G5   partialgpool_5.dense_part.0
    for (int m = 0; m < 4; m++) { add(4); add(4); add(1); }   // head: output_layers_PI / _V folded, 4 M tiles
    p.nt = t;
    p.total_bytes = o;
    return p;
}

struct Smem {
    __align__(128) unsigned char oper[OPER_BYTES];
    __align__(128) unsigned char pool[POOL_BYTES];
    __align__(128) unsigned char ring[RING][TILE];
    uint64_t full[RING], empty[RING];   // weight tiles: producer -> MMA issuer -> slot free
    uint64_t acc[2];                    // tcgen05.commit: the accumulator of half h of a 2d layer is complete
    uint64_t acc1d;                     // two arrivals (tcgen05.commit of both issuers): the accumulators of a 1d layer are complete
    uint64_t epi[2];                    // 16 arrivals (one per epilogue warp): the B operand of half h (1d layers: epi[0]) is written
    uint32_t tmem_base;
    uint64_t inbar;                     // the staged input rows have landed
    __align__(16) uint8_t rowsrc[NL];   // staged copy of row_src for the CTA's rows
    uint8_t rowflag[NL + 8];            // per row of the CTA: 0 beyond n_rows, 1 states / valids, 2 staging row
};
static_assert(sizeof(Smem) <= 227 * 1024, "shared memory budget");

__device__ long long g_stamps[32];      // diagnostics: phase time stamps of CTA 0 (SM clock)
__device__ long long g_tile_stamps[3][MAX_TILES];   // diagnostics, CTA 0: tile requested by the producer / landed (seen by the MMA issuer) / its MMAs issued
__device__ int g_err;                   // diagnostics: an mbarrier wait ran into its poll limit (a protocol error, never expected)
#define NN2_STAMP(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && threadIdx.x < 32) g_stamps[i] = clock64(); } while (0)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
// one lane of the (converged) warp; the compiler keeps the operands of what follows in uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0u;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// tcgen05.ld 32x32b: N consecutive fp32 columns of this thread's TMEM lane (no wait)
__device__ __forceinline__ void tld8(uint32_t a, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tld16(uint32_t a, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a) : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tld32(uint32_t a, float* v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(a) : "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack2(float x, float y) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int NP>
__global__ void __launch_bounds__(THREADS, 1) nnet2_forward_kernel(const unsigned char* __restrict__ blob, Plan plan, const int8_t* __restrict__ states,
                                                                   const uint8_t* __restrict__ valids, const uint8_t* __restrict__ row_src,
                                                                   const int8_t* __restrict__ alt_states, int alt_stride,
                                                                   const uint32_t* __restrict__ alt_mask, int alt_mask_stride, int n_rows,
                                                                   float* __restrict__ pi, float* __restrict__ vout) {
    constexpr int R = 32 + 10 * NP + NP * NP, S = 7 * R, K1 = (R + 15) / 16 * 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const float* __restrict__ prm = reinterpret_cast<const float*>(blob);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int base = blockIdx.x * NL;
    bool dead = false;                   // a wait that ran into its poll limit: record it and stop waiting (results are then garbage, nothing hangs)
    auto wait = [&](uint64_t* bar, uint32_t parity) {
        if (!dead && !umma::mbar_wait(bar, parity, 1u << 20)) { dead = true; atomicExch(&g_err, 1 + (int)(bar - sm.full)); }
    };
    NN2_STAMP(0);

    if (warp == MMA_WARP) umma::tmem_alloc(&sm.tmem_base, 512);
    if (tid == EPI_THREADS) {
        for (int i = 0; i < RING; i++) { umma::mbar_init(&sm.full[i], 1); umma::mbar_init(&sm.empty[i], 1); }
        umma::mbar_init(&sm.acc[0], 1); umma::mbar_init(&sm.acc[1], 1); umma::mbar_init(&sm.inbar, 1); umma::mbar_init(&sm.acc1d, 2);
        umma::mbar_init(&sm.epi[0], EPI_THREADS / 32); umma::mbar_init(&sm.epi[1], EPI_THREADS / 32);
    }
    if (tid < EPI_THREADS) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < POOL_BYTES / 16; i += EPI_THREADS) reinterpret_cast<uint4*>(sm.pool)[i] = z;
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = sm.tmem_base;
    NN2_STAMP(13);

    if (warp >= 16 && warp < MMA_WARP) {
        // ---------------------------------------------------------------- weight producers: tile t -> slot t % RING, requested by warp 16 + t % 3.
        // The whole warp runs the loop; one elected lane issues the copy
        for (int t = warp - 16; t < plan.nt; t += PRODUCERS) {
            const int sl = t % RING;
            if (t >= RING) { if (elect_one()) wait(&sm.empty[sl], (uint32_t)((t / RING - 1) & 1)); __syncwarp(); }
            if (elect_one()) {
                umma::mbar_expect(&sm.full[sl], (uint32_t)plan.bytes[t]);
                umma::bulk_load(sm.ring[sl], blob + plan.off[t], (uint32_t)plan.bytes[t], &sm.full[sl]);
                if (blockIdx.x == 0) g_tile_stamps[0][t] = clock64();
            }
            __syncwarp();
        }
    } else if (warp == MMA_WARP || warp == MMA_WARP_B) {
        // ---------------------------------------------------------------- MMA issuer: the whole warp runs the (uniform) control flow, the
        // tcgen05.mma / tcgen05.commit instructions themselves are issued by one elected lane (operands stay in uniform registers; a
        // single divergent lane made every MMA a ~25-instruction loop that could not keep up next to four busy epilogue warps)
        const uint32_t oper = umma::smem_u32(sm.oper), pool = umma::smem_u32(sm.pool);
        const uint32_t id224 = umma::instr_desc_bf16_bmn(128, HC), id64 = umma::instr_desc_bf16_bmn(128, NL);
        const int who = warp - MMA_WARP;           // issuer A (0) or B (1)
        uint32_t epi_n[2] = {0u, 0u};
        auto commit1 = [&](uint64_t* bar) { if (elect_one()) umma::commit(bar); __syncwarp(); };
        // the MMAs of one product: tiles [t0, t0 + nd) against the operand at b_addr (consecutive k), then (pooled) one K = 16 step
        // against the side operand. wait_full: first use of these tiles; free: give each slot back as soon as its MMAs are done
        // split: the tiles of the product alternate between the two issuers (tile j belongs to issuer j & 1; each issuer accumulates into
        // its own columns `dcol` and the epilogue adds the two accumulators)
        auto product = [&](int t0, int nd, bool pooled, uint32_t b_addr, uint32_t b_sbo, uint32_t p_addr, uint32_t idesc, uint32_t dcol,
                           bool wait_full, bool free_slots, bool split = false) {
            bool accum = false;
            for (int j = 0; j < nd + (pooled ? 1 : 0); j++) {
                if (split && (j & 1) != who) { if (j < nd) b_addr += (uint32_t)(plan.bytes[t0 + j] / KSTEP_BYTES) * 256; continue; }
                const int t = t0 + j, sl = t % RING;
                if (wait_full) { if (elect_one()) wait(&sm.full[sl], (uint32_t)((t / RING) & 1)); __syncwarp(); umma::fence_after_sync(); if (blockIdx.x == 0 && lane == 0) g_tile_stamps[1][t] = clock64(); }
                const int ks = plan.bytes[t] / KSTEP_BYTES;
                const bool side = pooled && j == nd;
                const uint64_t ad = umma::smem_desc(umma::smem_u32(sm.ring[sl]), 128, (uint32_t)ks * 256);
                const uint64_t bd = side ? umma::smem_desc(p_addr, 128, 256) : umma::smem_desc(b_addr, 128, b_sbo);
                if (elect_one()) {
                    for (int k = 0; k < ks; k++) { umma::mma_bf16_ss(tb + dcol, ad + (uint64_t)(16 * k), bd + (uint64_t)(16 * k), idesc, accum); accum = true; }
                    if (free_slots) umma::commit(&sm.empty[sl]);
                }
                __syncwarp();
                accum = true;
                if (!side) b_addr += (uint32_t)ks * 256;
                if (blockIdx.x == 0 && wait_full && lane == 0) g_tile_stamps[2][t] = clock64();
            }
        };
        auto wait_epi = [&](int h) { if (elect_one()) wait(&sm.epi[h], epi_n[h] & 1u); __syncwarp(); epi_n[h]++; umma::fence_after_sync(); };
        int t0 = 0;
        // 2d layers: L1, L2, G1, L3; the two halves share every tile. Issuer A does them; issuer B only follows the barrier phases
#pragma unroll 1
        for (int l = 0; l < 4; l++) {
            const int nd = l == 0 ? plan.nt_l1 : 2, ntl = nd + (l == 3 ? 1 : 0);
            const uint32_t koff = l == 2 ? 4u * 128u : 0u;          // G1 reads features 32..127
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                wait_epi(h);
                if (who == 0) {
                    product(t0, nd, l == 3, oper + h * HALF_BYTES + koff, 2048, pool + h * (HC / 8) * 256, id224, 256u * h, h == 0, false);
                    commit1(&sm.acc[h]);
                }
            }
            if (who == 0)
                for (int j = 0; j < ntl; j++) commit1(&sm.empty[(t0 + j) % RING]);
            t0 += ntl;
        }
        // 1d layers, N = 64 leaves: the weight tiles of a layer alternate between the two issuers, each accumulates its share of K into its
        // own 64 columns (A: 0..63, B: 64..127), the epilogue adds the two; both commits arrive on acc1d
        const uint32_t dc = 64u * (uint32_t)who;
        wait_epi(0); wait_epi(1);                                                                                     // L4: both halves of the 704-row operand
        product(t0, 11, false, oper, FLAT_SBO, 0u, id64, dc, true, true, true); commit1(&sm.acc1d); t0 += 11;
        wait_epi(0); product(t0, 2, false, oper + VEC_OFF + 2 * 128, 2048, 0u, id64, dc, true, true, true); commit1(&sm.acc1d); t0 += 2;   // G4
        wait_epi(0); product(t0, 2, true, oper + VEC_OFF, 2048, pool, id64, dc, true, true, true); commit1(&sm.acc1d); t0 += 3;          // L5a
        wait_epi(0); product(t0, 2, false, oper + VEC_OFF, 2048, 0u, id64, dc, true, true, true); commit1(&sm.acc1d); t0 += 2;           // L5b
        wait_epi(0); product(t0, 2, false, oper + VEC_OFF + 2 * 128, 2048, 0u, id64, dc, true, true, true); commit1(&sm.acc1d); t0 += 2;   // G5
        // head: 4 M tiles of 3 weight tiles each, stored in the order 0, 2, 1, 3: A takes M tiles 0 and 1, B takes 2 and 3 (whole products,
        // columns 64 m as before)
        wait_epi(0);
#pragma unroll 1
        for (int i = 0; i < 2; i++) product(t0 + 6 * i + 3 * who, 2, true, oper + VEC_OFF, 2048, pool, id64, 64u * (uint32_t)(2 * who + i), true, true);
        commit1(&sm.acc1d);
    } else {
        // ---------------------------------------------------------------- epilogue warps: thread = TMEM lane f (output feature), column group cg
        const int q = warp & 3, cg = warp >> 2, f = 32 * q + lane;
        const uint32_t lane_addr = tb + ((uint32_t)(32 * q) << 16);
        // programmatic dependent launch: everything above may run while the kernel that produces the input rows is still draining
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        // ---- input. Operand element (k, column 7 s + c) = state[s][k][c] (int8 counts are exact in bf16), zero up to K1; one thread per
        // (k, 8 consecutive columns): the 8 bytes come from at most two leaves, consecutive threads take consecutive k.
        // rowflag: where each of the CTA's 64 rows comes from (0: beyond n_rows, 1: states / valids, 2: the tree arena's staging row).
        constexpr int CH = K1 * (HC / 8), IT = (CH + EPI_THREADS - 1) / EPI_THREADS;
        constexpr int STG_ALT = (NL * S + 511) / 512 * 512;      // staging: the 64 rows of `states`, then (from here) the 64 staging rows
        const bool staged = base + NL <= n_rows && (!row_src || STG_ALT + NL * alt_stride <= HALF_BYTES) && STG_ALT <= HALF_BYTES &&
                            (NL * S) % 16 == 0 && (NL * alt_stride) % 16 == 0 &&
                            (((uintptr_t)states | (uintptr_t)alt_states | (uintptr_t)row_src) & 15u) == 0;
        if (staged) {
            // full CTA: three bulk copies (rows of both sources + the 64 source flags), issued by three different warps so that they
            // travel at the same time, into the second half of the operand region; conversion reads shared memory
            unsigned char* stg = sm.oper + HALF_BYTES;
            if (warp < 3 && elect_one()) {
                if (warp == 0) {
                    umma::mbar_expect(&sm.inbar, (uint32_t)(NL * S) + (row_src ? (uint32_t)(NL * alt_stride) + NL : 0u));
                    umma::bulk_load(stg, states + (size_t)base * S, (uint32_t)(NL * S), &sm.inbar);
                } else if (row_src && warp == 1) {
                    umma::bulk_load(stg + STG_ALT, alt_states + (size_t)base * alt_stride, (uint32_t)(NL * alt_stride), &sm.inbar);
                } else if (row_src && warp == 2) {
                    umma::bulk_load(sm.rowsrc, row_src + base, NL, &sm.inbar);
                }
            }
            __syncwarp();
            if (lane == 0) wait(&sm.inbar, 0u);
            __syncwarp();
            if (tid < NL) sm.rowflag[tid] = (uint8_t)((row_src && sm.rowsrc[tid]) ? 2 : 1);
            epi_bar_sync();
            NN2_STAMP(14);
            auto chunk = [&](int h, int i, uint32_t (&pk)[4]) {
                const int n8 = i / K1, k = i - n8 * K1, s0 = (8 * n8) / 7;
                // elements e < cnt0 belong to leaf s0 (gem column c0 + e), the others to leaf s0 + 1 (gem column e - cnt0)
                const int c0 = 8 * n8 - 7 * s0, cnt0 = 7 - c0;
                const int sl0 = HL * h + s0, sl1 = min(sl0 + 1, NL - 1);
                const signed char* q0 = reinterpret_cast<const signed char*>(sm.rowflag[sl0] == 2 ? stg + STG_ALT + sl0 * alt_stride : stg + sl0 * S) + k * 7 + c0;
                const signed char* q1 = reinterpret_cast<const signed char*>(sm.rowflag[sl1] == 2 ? stg + STG_ALT + sl1 * alt_stride : stg + sl1 * S) + k * 7 - cnt0;
                if (k < R) {
                    float x[8];
#pragma unroll
                    for (int e = 0; e < 8; e++)      // int -> float without the conversion pipe: 1.5 * 2^23 + v is exact in the mantissa
                        x[e] = __int_as_float(0x4B400000 + (int)(e < cnt0 ? q0 : q1)[e]) - 12582912.f;
#pragma unroll
                    for (int e = 0; e < 8; e += 2) pk[e >> 1] = pack2(x[e], x[e + 1]);
                } else {
                    pk[0] = pk[1] = pk[2] = pk[3] = 0u;
                }
            };
#pragma unroll
            for (int it = 0; it < IT; it++) {
                const int i = tid + it * EPI_THREADS;
                if (i < CH) {
                    uint32_t pk[4];
                    chunk(0, i, pk);
                    *reinterpret_cast<uint4*>(sm.oper + (i / K1) * 2048 + (i % K1) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            umma::fence_smem_to_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.epi[0]);
            NN2_STAMP(15);
            // second half: its operand overwrites the staging area, so everything is read before anything is written
            uint32_t pk[IT][4];
#pragma unroll
            for (int it = 0; it < IT; it++) {
                const int i = tid + it * EPI_THREADS;
                if (i < CH) chunk(1, i, pk[it]);
            }
            epi_bar_sync();
#pragma unroll
            for (int it = 0; it < IT; it++) {
                const int i = tid + it * EPI_THREADS;
                if (i < CH) *reinterpret_cast<uint4*>(sm.oper + HALF_BYTES + (i / K1) * 2048 + (i % K1) * 16) = make_uint4(pk[it][0], pk[it][1], pk[it][2], pk[it][3]);
            }
            umma::fence_smem_to_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.epi[1]);
        } else {
            // last (partial) CTA, unaligned buffers: plain loads
            if (tid < NL) {
                const int row = base + tid;
                sm.rowflag[tid] = (uint8_t)(row < n_rows ? ((row_src && row_src[row]) ? 2 : 1) : 0);
            }
            epi_bar_sync();
            NN2_STAMP(14);
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
#pragma unroll 1
                for (int i = tid; i < CH; i += EPI_THREADS) {
                    const int n8 = i / K1, k = i - n8 * K1, s0 = (8 * n8) / 7;
                    const int8_t* p0 = nullptr;
                    const int8_t* p1 = nullptr;
                    {
                        const int sl = HL * h + s0, fl = sm.rowflag[sl];
                        if (k < R && fl) p0 = (fl == 2 ? alt_states + (size_t)(base + sl) * alt_stride : states + (size_t)(base + sl) * S) + k * 7;
                    }
                    if (s0 + 1 < HL) {
                        const int sl = HL * h + s0 + 1, fl = sm.rowflag[sl];
                        if (k < R && fl) p1 = (fl == 2 ? alt_states + (size_t)(base + sl) * alt_stride : states + (size_t)(base + sl) * S) + k * 7;
                    }
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        float x[2];
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            const int n = 8 * n8 + e + u, sx = n / 7, c = n - 7 * sx;
                            const int8_t* src = sx == s0 ? p0 : p1;
                            x[u] = src ? (float)src[c] : 0.f;
                        }
                        pk[e >> 1] = pack2(x[0], x[1]);
                    }
                    *reinterpret_cast<uint4*>(sm.oper + h * HALF_BYTES + n8 * 2048 + k * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                umma::fence_smem_to_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.epi[h]);
            }
        }
        NN2_STAMP(1);
        uint32_t acc_n[2] = {0u, 0u};
        // one lane polls, one lane arrives: 512 threads on one mbarrier word serialise (measured: ~1.3 us per arrival round)
        auto wait_acc = [&](int h) {
            if (lane == 0) wait(&sm.acc[h], acc_n[h] & 1u);
            acc_n[h]++;
            __syncwarp();
            umma::fence_after_sync();
        };
        auto publish = [&](int h) {
            umma::fence_smem_to_async();
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.epi[h]);
        };

        // per-thread constants of the 2d layers, fetched once: bias of this thread's feature and, per gem column c, the folded
        // BatchNorm1d(7) terms: out = relu(acc * sc[c] + (b * sc[c] + sh[c])) is ONE fma per element; rows the layer does not have (G1:
        // features 120..127) get scale = shift = 0 and come out as the zeros the next product expects
        const float b_l1 = __ldg(prm + P_B1 + f), b_l2 = __ldg(prm + P_B2 + f), b_g1 = __ldg(prm + P_BG1 + f), b_l3 = __ldg(prm + P_B3 + f);
        // relu on the packed bf16 pair (max with +0 after rounding = rounding after max)
        auto relu2 = [](float x, float y) -> uint32_t {
            const __nv_bfloat162 r = __hmax2(__floats2bfloat162_rn(x, y), __floats2bfloat162_rn(0.f, 0.f));
            return *reinterpret_cast<const uint32_t*>(&r);
        };
        // max / mean over `rows` consecutive 16-byte rows (8 columns each) of the bf16 operand -> rows g and 4 + g of the side operand
        auto pool_rows = [&](const unsigned char* src, int rows, float inv, unsigned char* dp, int g) {
            float mx[8], sum[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { mx[i] = -INFINITY; sum[i] = 0.f; }
            for (int r = 0; r < rows; r++) {
                const uint4 w4 = reinterpret_cast<const uint4*>(src)[r];
                const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                    mx[2 * i] = fmaxf(mx[2 * i], x.x); mx[2 * i + 1] = fmaxf(mx[2 * i + 1], x.y);
                    sum[2 * i] += x.x; sum[2 * i + 1] += x.y;
                }
            }
            *reinterpret_cast<uint4*>(dp + g * 16) = make_uint4(pack2(mx[0], mx[1]), pack2(mx[2], mx[3]), pack2(mx[4], mx[5]), pack2(mx[6], mx[7]));
            *reinterpret_cast<uint4*>(dp + (4 + g) * 16) = make_uint4(pack2(sum[0] * inv, sum[1] * inv), pack2(sum[2] * inv, sum[3] * inv),
                                                                        pack2(sum[4] * inv, sum[5] * inv), pack2(sum[6] * inv, sum[7] * inv));
        };

        // ---- 2d layers
#pragma unroll 1
        for (int l = 0; l < 4; l++) {
            float sc[7], bs[7];
            {
                const bool bn = l == 0 || l == 2, live = !(l == 2 && f >= 120);
                const float b = l == 0 ? b_l1 : l == 1 ? b_l2 : l == 2 ? b_g1 : b_l3;
#pragma unroll
                for (int c = 0; c < 7; c++) {
                    const float s_ = bn ? __ldg(prm + (l == 0 ? P_S1 : P_SG1) + c) : 1.f, t_ = bn ? __ldg(prm + (l == 0 ? P_T1 : P_TG1) + c) : 0.f;
                    sc[c] = live ? s_ : 0.f;
                    bs[c] = live ? fmaf(b, s_, t_) : 0.f;
                }
            }
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                wait_acc(h);
                if (l == 2) NN2_STAMP(16 + 4 * h);
                float v[56];
                const uint32_t ta = lane_addr + 256u * h + 56u * cg;
                tld32(ta, v); tld16(ta + 32, v + 32); tld8(ta + 48, v + 48);
                tld_wait();
                if (l == 2) NN2_STAMP(17 + 4 * h);
                if (l < 3) {
                    unsigned char* dst = sm.oper + h * HALF_BYTES + (7 * cg) * 2048 + f * 16;
#pragma unroll
                    for (int j8 = 0; j8 < 7; j8++) {
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            const int j = 8 * j8 + e;
                            pk[e >> 1] = relu2(fmaf(v[j], sc[j % 7], bs[j % 7]), fmaf(v[j + 1], sc[(j + 1) % 7], bs[(j + 1) % 7]));
                        }
                        *reinterpret_cast<uint4*>(dst + j8 * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                    if (l == 2) NN2_STAMP(18 + 4 * h);
                    publish(h);
                    if (l == 2) NN2_STAMP(19 + 4 * h);
                    if (l == 1 && q == 0) {
                        // the pooled half of DenseAndPartialGPool(4 x 8): max and mean over features 8 g .. 8 g + 7 of the bf16 activations, per
                        // column. Features 0..31 of this warp's 56 columns were written by this very warp (q = 0: f = lane), and the next layer's
                        // epilogue that overwrites them is this warp's too: no CTA barrier, 28 lanes take one (8 columns, group) task each - the
                        // 8 x 16 bytes of a task are one core matrix. The next publish of this warp makes the side operand visible to its reader
                        __syncwarp();
                        if (lane < 28) {
                            const int n8 = 7 * cg + (lane >> 2), g = lane & 3;
                            pool_rows(sm.oper + h * HALF_BYTES + n8 * 2048 + g * 128, 8, 0.125f, sm.pool + (h * (HC / 8) + n8) * 256, g);
                        }
                        __syncwarp();
                    }
                } else {
                    // L3 + FlattenAndPartialGPool(64, 5): this thread holds feature f of the 7 gem columns of 8 leaves (one leaf octet of the
                    // 704-row operand): rows [max over the 5 colours | mean | gold | points] for f < 64, the 7 columns as they are for f >= 64
                    unsigned char* dst = sm.oper + (4 * h + cg) * FLAT_SBO;
                    if (f < 64) {
                        float o[4][8];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            float a[7];
#pragma unroll
                            for (int c = 0; c < 7; c++) a[c] = fmaxf(v[7 * i + c] + b_l3, 0.f);
                            o[0][i] = fmaxf(fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3])), a[4]);
                            o[1][i] = (a[0] + a[1] + a[2] + a[3] + a[4]) * 0.2f;
                            o[2][i] = a[5]; o[3][i] = a[6];
                        }
#pragma unroll
                        for (int r = 0; r < 4; r++)
                            *reinterpret_cast<uint4*>(dst + (64 * r + f) * 16) =
                                make_uint4(pack2(o[r][0], o[r][1]), pack2(o[r][2], o[r][3]), pack2(o[r][4], o[r][5]), pack2(o[r][6], o[r][7]));
                    } else {
#pragma unroll
                        for (int c = 0; c < 7; c++)
                            *reinterpret_cast<uint4*>(dst + (256 + 64 * c + (f - 64)) * 16) =
                                make_uint4(relu2(v[c] + b_l3, v[7 + c] + b_l3), relu2(v[14 + c] + b_l3, v[21 + c] + b_l3),
                                           relu2(v[28 + c] + b_l3, v[35 + c] + b_l3), relu2(v[42 + c] + b_l3, v[49 + c] + b_l3));
                    }
                    publish(h);
                }
                if (h == 1) NN2_STAMP(2 + l);
            }
        }

        // ---- 1d layers: columns = the 64 leaves, this thread finishes feature f of leaves 16 cg .. 16 cg + 15
        unsigned char* vec = sm.oper + VEC_OFF;
        float b1d[5];
#pragma unroll
        for (int l = 0; l < 5; l++) b1d[l] = __ldg(prm + (l == 0 ? P_B4 : l == 1 ? P_BG4 : l == 2 ? P_B5A : l == 3 ? P_B5B : P_BG5) + f);
        float bh[4];
#pragma unroll
        for (int m = 0; m < 4; m++) bh[m] = __ldg(prm + P_BH + 128 * m + f);
        // legality of the four rows this warp finishes in the softmax (bit j of vbits[r] = action lane + 32 j of row 4 warp + r): requested here,
        // where registers are free, and turned into bits after the wait for the 704-wide layer - the loads travel under that wait
        uint32_t vraw[4][13], vbits[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int fl = sm.rowflag[4 * warp + r];
            const size_t row = (size_t)(base + 4 * warp + r);
#pragma unroll
            for (int j = 0; j < 13; j++) {
                const int a = lane + 32 * j;
                vraw[r][j] = fl == 2 ? alt_mask[(size_t)j * alt_mask_stride + row] : (fl == 1 && a < NN_ACTIONS) ? (uint32_t)valids[row * NN_ACTIONS + a] : 0u;
            }
        }
        uint32_t acc1d_n = 0u;
        auto wait_acc1d = [&]() {
            if (lane == 0) wait(&sm.acc1d, acc1d_n & 1u);
            acc1d_n++;
            __syncwarp();
            umma::fence_after_sync();
        };
#pragma unroll
        for (int l = 0; l < 5; l++) {      // L4, G4, L5a, L5b, G5
            wait_acc1d();
            if (l == 0) {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const bool words = sm.rowflag[4 * warp + r] == 2;
#pragma unroll
                    for (int j = 0; j < 13; j++) vbits[r] |= (words ? ((vraw[r][j] >> lane) & 1u) : (vraw[r][j] != 0u ? 1u : 0u)) << j;
                }
            }
            float v[16], v2[16];
            tld16(lane_addr + 16u * cg, v);                 // issuer A's share of K
            tld16(lane_addr + 64u + 16u * cg, v2);          // issuer B's
            tld_wait();
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] += v2[j];
            const bool live = !((l == 1 || l == 4) && f >= 120);
            const float b = b1d[l];
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) pk[j >> 1] = live ? relu2(v[j] + b, v[j + 1] + b) : 0u;
            *reinterpret_cast<uint4*>(vec + (2 * cg) * 2048 + f * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(vec + (2 * cg + 1) * 2048 + f * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            publish(0);
            if ((l == 0 || l == 3) && q == 0) {
                // DenseAndPartialGPool(4 x 4): max and mean over features 4 g .. 4 g + 3 (written by lanes 0..15 of this warp) of this
                // warp's 16 leaves -> rows g and 4 + g of the side operand; 8 lanes, one (8 leaves, group) task each
                __syncwarp();
                if (lane < 8) {
                    const int n8 = 2 * cg + (lane >> 2), g = lane & 3;
                    pool_rows(vec + n8 * 2048 + g * 64, 4, 0.25f, sm.pool + n8 * 256, g);
                }
                __syncwarp();
            }
            NN2_STAMP(6 + l);
        }

        // ---- head: logits of action 128 m + f (+ bias) as fp32 rows in shared memory; the value rows 406 .. 406 + n - 1 go out as tanh
        {
            wait_acc1d();
            float* logits = reinterpret_cast<float*>(sm.oper);
            float v[64];
#pragma unroll
            for (int m = 0; m < 4; m++) tld16(lane_addr + 64u * m + 16u * cg, v + 16 * m);
            tld_wait();
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int a = 128 * m + f;
                if (a < NN_ACTIONS) {
#pragma unroll
                    for (int j = 0; j < 16; j++) logits[(16 * cg + j) * LSTR + a] = v[16 * m + j] + bh[m];
                } else if (a < NN_ACTIONS + NP) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int row = base + 16 * cg + j;
                        if (row < n_rows) vout[(size_t)row * NP + (a - NN_ACTIONS)] = tanhf(v[16 * m + j] + bh[m]);
                    }
                }
            }
            umma::fence_before_sync();
            epi_bar_sync();
            NN2_STAMP(11);
            // ---- masked softmax: log_softmax(where(valid, pi, -1e8)) then exp (SplendorNNet.py:153-159, GenericNNetWrapper.py:166)
#pragma unroll 1
            for (int r0 = 0; r0 < 4; r0 += 2) {      // two rows at a time: their load / reduce / exp chains interleave
                float x[2][13], mx[2], sum[2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int s = 4 * warp + r0 + u;
                    const uint32_t vb = vbits[r0 + u];
                    mx[u] = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 13; j++) {
                        const int a = lane + 32 * j;
                        x[u][j] = -INFINITY;
                        if (a < NN_ACTIONS) {
                            x[u][j] = ((vb >> j) & 1u) ? logits[s * LSTR + a] : -1e8f;
                            mx[u] = fmaxf(mx[u], x[u][j]);
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    mx[0] = fmaxf(mx[0], __shfl_xor_sync(0xffffffffu, mx[0], o));
                    mx[1] = fmaxf(mx[1], __shfl_xor_sync(0xffffffffu, mx[1], o));
                }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    sum[u] = 0.f;
#pragma unroll
                    for (int j = 0; j < 13; j++) {
                        x[u][j] = (lane + 32 * j) < NN_ACTIONS ? __expf(x[u][j] - mx[u]) : 0.f;
                        sum[u] += x[u][j];
                    }
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], o);
                    sum[1] += __shfl_xor_sync(0xffffffffu, sum[1], o);
                }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int s = 4 * warp + r0 + u;
                    if (base + s >= n_rows) continue;
                    const float inv = 1.f / sum[u];
#pragma unroll
                    for (int j = 0; j < 13; j++) {
                        const int a = lane + 32 * j;
                        if (a < NN_ACTIONS) pi[(size_t)(base + s) * NN_ACTIONS + a] = x[u][j] * inv;
                    }
                }
            }
            NN2_STAMP(12);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) umma::tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------ host: folding + blob packing
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
// one weight tile [128 rows][16 ks columns] in the canonical K-major layout: element (r, k) at (r / 8) * (2 ks) * 128 + (k / 8) * 128 +
// (r % 8) * 16 + (k % 8) * 2; `at(r, k)` supplies the value (0 outside the matrix)
template <class F>
void pack_tile(unsigned char* dst, int ks, F at) {
    uint16_t* d = reinterpret_cast<uint16_t*>(dst);
    const int kc = 2 * ks;
    for (int r = 0; r < 128; r++)
        for (int k = 0; k < 16 * ks; k++)
            d[((size_t)(r >> 3) * kc * 128 + (size_t)(k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2) / 2] = to_bf16(at(r, k));
}

size_t blob_bytes(int n) { return (size_t)make_plan(n).total_bytes; }

int pack(int n, const float* const* T, void* blob, size_t blob_bytes_) {
    const Plan p = make_plan(n);
    if (blob_bytes_ < (size_t)p.total_bytes) return spl_fail_(SPL_E_ARG, "spl_nnet_pack: blob smaller than spl_nnet_blob_bytes");
    const int R = 32 + 10 * n + n * n;
    unsigned char* B = (unsigned char*)blob;
    memset(B, 0, p.total_bytes);
    float* prm = reinterpret_cast<float*>(B);
    const BnFold bn1 = fold_bn(T[2], T[3], T[4], T[5], 7), bng1 = fold_bn(T[10], T[11], T[12], T[13], 7);
    const BnFold bn4 = fold_bn(T[20], T[21], T[22], T[23], 1), bn5 = fold_bn(T[26], T[27], T[28], T[29], 1), bng5 = fold_bn(T[34], T[35], T[36], T[37], 1);
    for (int i = 0; i < 128; i++) {
        prm[P_B1 + i] = T[1][i]; prm[P_B2 + i] = T[7][i]; prm[P_B3 + i] = T[15][i]; prm[P_B4 + i] = T[17][i];
        prm[P_B5A + i] = (float)((double)T[25][i] * bn5.s[0] + bn5.t[0]);
        prm[P_B5B + i] = T[31][i];
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
    // head: output_layers_PI.1 . output_layers_PI.0 (406 x 128) and output_layers_V.1 . output_layers_V.0 (n x 128), biases likewise
    std::vector<double> WH((size_t)512 * 128, 0.0), BH(512, 0.0);
    auto fold_head = [&](const float* W0, const float* b0, const float* W1, const float* b1, int rows, int row0) {
        for (int r = 0; r < rows; r++) {
            double bb = (double)b1[r];
            for (int j = 0; j < 128; j++) bb += (double)W1[(size_t)r * 128 + j] * (double)b0[j];
            BH[row0 + r] = bb;
            for (int i = 0; i < 128; i++) {
                double s = 0.0;
                for (int j = 0; j < 128; j++) s += (double)W1[(size_t)r * 128 + j] * (double)W0[(size_t)j * 128 + i];
                WH[(size_t)(row0 + r) * 128 + i] = s;
            }
        }
    };
    fold_head(T[38], T[39], T[40], T[41], NN_ACTIONS, 0);
    fold_head(T[42], T[43], T[44], T[45], n, NN_ACTIONS);
    for (int i = 0; i < 512; i++) prm[P_BH + i] = (float)BH[i];

    int t = 0;
    auto tile = [&](auto at) { pack_tile(B + p.off[t], p.bytes[t] / KSTEP_BYTES, at); t++; };
    // plain [rows][K] matrix W (row-major, `rows` real rows) against operand features k0 .. : tiles of 64 k
    auto dense = [&](const float* W, int rows, int K, int kfirst, int kcount, double scale) {
        for (int k0 = 0; k0 < kcount; k0 += 64)
            tile([&](int r, int k) { const int kk = k0 + k; return (r < rows && kk < kcount) ? (double)W[(size_t)r * K + kfirst + kk] * scale : 0.0; });
    };
    // a layer whose input is [4 max | 4 mean | 120 dense]: operand rows 0..119 = the dense inputs (columns 8..127 of W), then the side step
    auto after_pool = [&](auto w_at, int rows) {
        for (int k0 = 0; k0 < 128; k0 += 64)
            tile([&](int r, int k) { const int kk = k0 + k; return (r < rows && kk < 120) ? w_at(r, 8 + kk) : 0.0; });
        tile([&](int r, int k) { return (r < rows && k < 8) ? w_at(r, k) : 0.0; });
    };
    dense(T[0], 128, R, 0, R, 1.0);                                     // L1: K1 = 16-padded R
    dense(T[6], 128, 128, 0, 128, 1.0);                                 // L2
    dense(T[8], 120, 96, 0, 96, 1.0);                                   // G1 (BatchNorm1d(7) in the epilogue)
    after_pool([&](int r, int k) { return (double)T[14][(size_t)r * 128 + k]; }, 128);                 // L3
    dense(T[16], 128, 704, 0, 704, 1.0);                                // L4
    dense(T[18], 120, 112, 0, 112, bn4.s[0]);                           // G4
    after_pool([&](int r, int k) { return (double)T[24][(size_t)r * 128 + k] * bn5.s[0]; }, 128);      // L5a
    dense(T[30], 128, 128, 0, 128, 1.0);                                // L5b
    dense(T[32], 120, 112, 0, 112, bng5.s[0]);                          // G5
    for (int m : {0, 2, 1, 3})       // head: M tiles 0 and 1 belong to issuer A, 2 and 3 to issuer B; stored in the order they are consumed
        after_pool([&](int r, int k) { return WH[(size_t)(128 * m + r) * 128 + k]; }, 128);
    if (t != p.nt) return spl_fail_(SPL_E_ARG, "spl_nnet_pack: internal tile count mismatch");
    return SPL_OK;
}

int debug_stamps(long long* out32) {
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out32, g_stamps, sizeof(long long) * 32));
    int e = 0;
    CU(cudaMemcpyFromSymbol(&e, g_err, sizeof(int)));
    out32[31] = e;
    return SPL_OK;
}

int debug_tile_stamps(long long* out) {   // [3][MAX_TILES]
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out, g_tile_stamps, sizeof(long long) * 3 * MAX_TILES));
    return SPL_OK;
}

int forward_rows(spl_ctx* c, const void* blob, const int8_t* states, const uint8_t* valids, const uint8_t* row_src, const int8_t* alt_states,
                 int alt_stride, const uint32_t* alt_mask, int alt_mask_stride, int n_rows, float* pi, float* v, cudaStream_t st,
                 bool programmatic_dependent) {
    const Plan p = make_plan(c->n);
    const int grid = (n_rows + NL - 1) / NL;
    const int smem = (int)sizeof(Smem);
    DISPATCH_N(c->n, {
        auto k = nnet2_forward_kernel<N>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cfg.attrs = attr; cfg.numAttrs = programmatic_dependent ? 1 : 0;
        CU(cudaLaunchKernelEx(&cfg, k, (const unsigned char*)blob, p, states, valids, row_src, alt_states, alt_stride, alt_mask, alt_mask_stride, n_rows, pi, v));
    });
    CU(cudaGetLastError());
    return SPL_OK;
}

}   // namespace nn2
