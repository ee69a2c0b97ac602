// libsplendor_b200.so - batched Splendor environment kernels for sm_100a and their C ABI
// (include/splendor_b200.h). Rules logic lives in spl_rules.cuh; this file is data movement,
// launch plumbing and the fused per-ply passes.
//
// Resident layout ("lane tiles"): int8[T][7R][32] - games are grouped in tiles of one warp (32 lanes);
// inside a tile the state is structure-of-arrays, cell-major, so (a) a tile is ONE contiguous
// 7R*32-byte block in HBM (12.25 / 15.5 / 19.25 KB for 2/3/4 players) that a single TMA bulk
// copy (cp.async.bulk, SASS UBLKCP) moves to shared memory and back, and (b) lane l of a warp
// reads byte [cell][l]: 32 consecutive bytes per warp access, no bank conflicts.
// One warp owns one tile: no block-wide barriers anywhere, every warp runs its own
// mbarrier-tracked load -> rules -> bulk store pipeline.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/splendor_b200.h"
#include "spl_rules.cuh"

#define TL SPL_LANE_TILE   // 32 lanes per tile
static_assert(TL == 32, "one warp per tile");

#include "spl_internal.h"

static thread_local char g_err[256] = "";
int spl_fail_(int code, const char* what, cudaError_t e) {
    if (e != cudaSuccess) snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof g_err, "%s", what);
    return code;
}
#define fail spl_fail_

// ------------------------------------------------------------------------------------------
// shared-memory tile accessor: cell (row, col) of this lane = byte [(7 row + col) * 32 + lane]
// ------------------------------------------------------------------------------------------
struct TileAcc {
    int8_t* b;   // tile base + lane
    __device__ __forceinline__ int get(int row, int col) const { return b[(row * 7 + col) * TL]; }
    __device__ __forceinline__ void set(int row, int col, int v) { b[(row * 7 + col) * TL] = (int8_t)v; }
};
// same cell order straight from a tile in global memory (read-mostly helpers: scores)
struct GTileAcc {
    const int8_t* b;
    __device__ __forceinline__ int get(int row, int col) const { return b[(row * 7 + col) * TL]; }
};

// ------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int BYTES, bool TMA>
__device__ __forceinline__ void tile_load(int8_t* sm, const int8_t* g, uint64_t* bar, uint32_t parity, int lane) {
    if (TMA) {
        if (lane == 0) {
            mbar_expect_tx(bar, BYTES);
            bulk_g2s(sm, g, BYTES, bar);
        }
        mbar_wait(bar, parity);
    } else {
        const int4* src = reinterpret_cast<const int4*>(g);
        int4* dst = reinterpret_cast<int4*>(sm);
#pragma unroll 4
        for (int i = lane; i < BYTES / 16; i += 32) dst[i] = __ldg(src + i);
        __syncwarp();
    }
}
template <int BYTES, bool TMA>
__device__ __forceinline__ void tile_store(int8_t* g, const int8_t* sm, int lane) {
    if (TMA) {
        fence_async_smem();   // generic-proxy writes of every lane -> visible to the async proxy
        __syncwarp();
        if (lane == 0) bulk_s2g(g, sm, BYTES);
    } else {
        __syncwarp();
        const int4* src = reinterpret_cast<const int4*>(sm);
        int4* dst = reinterpret_cast<int4*>(g);
#pragma unroll 4
        for (int i = lane; i < BYTES / 16; i += 32) dst[i] = src[i];
    }
}
// the tile buffer may be overwritten again only after the bulk store has read it
template <bool TMA>
__device__ __forceinline__ void tile_store_drain(int lane) {
    if (TMA) {
        if (lane == 0) bulk_wait_read();
        __syncwarp();
    } else {
        __syncwarp();
    }
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned v) { return (unsigned long long)__reduce_add_sync(0xffffffffu, v); }

// ------------------------------------------------------------------------------------------
// game start for some lanes of a tile, by the whole warp (init_game :222-246 with Philox chance). A finished game
// restarts in one lane at a time (one ply in three sees some lane of the warp finish), and the scalar spl_init_philox ran
// ~3000 instructions with a single active lane. Here the warp works on that lane's column together: the empty position is
// written 32 cells per instruction, the 14 Philox blocks (12 deals + nobles) are computed by 14 lanes at once, and the
// three tiers deal in parallel (the four draws of a tier depend on each other). Same counters, same result as
// spl_init_philox (the GPU tests compare every reset with the scalar form).
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void tile_reset_lanes(int8_t* tile, uint32_t need, int lane, uint64_t seed, uint32_t game, uint32_t episode) {
    typedef SplLay<N> L;
    while (need) {
        const int who = __ffs(need) - 1;
        need &= need - 1u;
        const uint32_t g = __shfl_sync(0xffffffffu, game, who), e = __shfl_sync(0xffffffffu, episode, who);
        int8_t* col = tile + who;
        for (int cell = lane; cell < L::CELLS; cell += 32) col[cell * TL] = (int8_t)spl_init_cell<N>(cell / 7, cell % 7);
        SplPhilox ph;
        ph.v[0] = ph.v[1] = ph.v[2] = ph.v[3] = 0u;
        if (lane < 12) ph = spl_philox(seed, g, e, (uint32_t)lane, 2);          // deal of visible slot `lane`
        else if (lane < 14) ph = spl_philox(seed, g, e, (uint32_t)(lane - 12), 3);   // nobles
        __syncwarp();
        TileAcc s{col};
#pragma unroll 1
        for (int round = 0; round < 4; round++) {   // lane t < 3 deals slot 4 t + round of tier t
            const int slot = 4 * (lane < 3 ? lane : 0) + round;
            const uint32_t w0 = __shfl_sync(0xffffffffu, ph.v[0], slot), w1 = __shfl_sync(0xffffffffu, ph.v[1], slot);
            if (lane < 3) {
                const int code = spl_draw_from<N>(s, lane, w0, w1);
                spl_write_card(s, L::CARDS + 2 * slot, spl_deck_take<N>(s, lane, code));
            }
            __syncwarp();
        }
        uint32_t w[5];
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = __shfl_sync(0xffffffffu, ph.v[i], 12);
        w[4] = __shfl_sync(0xffffffffu, ph.v[0], 13);
        if (lane == 0) spl_init_nobles<N>(s, w);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// fused single-ply pass (spl_step): one launch = make_move + swap_players + check_end_game
// (+ auto reset) + valid_moves (+ random pick) for every lane
// ------------------------------------------------------------------------------------------
struct StepParams {
    spl_step_args a;
    SplRules rules;
    int n_tiles;
};

template <int N, bool TMA, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) spl_step_kernel(const StepParams P) {
    typedef SplLay<N> L;
    constexpr int TB = L::CELLS * TL;
    extern __shared__ __align__(128) int8_t smem[];
    __shared__ uint64_t bars[WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * WARPS + warp;
    if (tile >= P.n_tiles) return;
    int8_t* sm = smem + (size_t)warp * TB;
    if (TMA) {
        if (lane == 0) mbar_init(&bars[warp], 1);
        __syncwarp();
    }
    const spl_step_args& A = P.a;
    const int gl = tile * TL + lane;
    const bool active = gl < A.n_lanes;
    int8_t* gtile = A.planes + (size_t)tile * TB;

    // per-lane scalars travel while the tile is in flight
    if (TMA && lane == 0) {
        mbar_expect_tx(&bars[warp], TB);
        bulk_g2s(sm, gtile, TB, &bars[warp]);
    }
    int p = A.player, a = -1, code = 255;
    uint32_t episode = 0;
    if (active) {
        if (A.players) p = A.players[gl];
        if (A.actions) a = A.actions[gl];
        if (A.reveals) code = A.reveals[gl];
        if (A.episodes) episode = A.episodes[gl];
    }
    if (TMA) mbar_wait(&bars[warp], 0);
    else tile_load<TB, false>(sm, gtile, nullptr, 0, lane);

    TileAcc s{sm + lane};
    const uint32_t game = A.game_base + (uint32_t)gl;
    int cur = p, status = p;
    unsigned moved = 0, finished = 0;
    float res[N];
    if (active) {
        if (a >= 0) {
            SplChance ch;
            ch.mode = A.chance_mode;
            if (ch.mode == SPL_CHANCE_REPLAY && code == 255) ch.mode = SPL_CHANCE_DETERMINISTIC;
            ch.code = code; ch.seed = A.seed; ch.game = game; ch.episode = episode;
            ch.ply = (uint32_t)(uint8_t)s.get(L::BANK, SPL_PTS);
            status = spl_apply_move<N>(s, a, p, ch);
            if (status >= 0) { cur = status; moved = 1; }
        }
        if (A.rotate) { spl_rotate<N>(s, cur, P.rules); cur = 0; }
        const bool ended = spl_game_ended<N>(s, P.rules, res);
        if (A.ended_out) {
#pragma unroll
            for (int i = 0; i < N; i++) A.ended_out[(size_t)gl * N + i] = res[i];
        }
        if (ended && A.auto_reset) {
            finished = 1;
            episode += 1;
            cur = 0;
            if (A.episodes) A.episodes[gl] = episode;
        }
        if (A.status_out) A.status_out[gl] = status;
    }
    {   // finished lanes start their next game: the whole warp deals for them, one lane at a time
        const uint32_t need = __ballot_sync(0xffffffffu, finished != 0);
        if (need) tile_reset_lanes<N>(sm, need, lane, A.seed, game, episode);
    }
    if (active) {
        if (A.mask_out || A.next_actions) {
            uint32_t m[SPL_MASK_WORDS];
            spl_valid_mask<N>(s, cur, P.rules, m);
            if (A.mask_out) {
                const size_t lp = (size_t)P.n_tiles * TL;
#pragma unroll
                for (int w = 0; w < SPL_MASK_WORDS; w++) A.mask_out[(size_t)w * lp + gl] = m[w];
            }
            if (A.next_actions)
                A.next_actions[gl] = (int16_t)spl_pick_random(m, A.seed, game, episode, (uint32_t)(uint8_t)s.get(L::BANK, SPL_PTS));
        }
    }
    if (A.counters) {
        unsigned long long f = warp_sum(finished), mv = warp_sum(moved);
        if (lane == 0) {
            if (f) atomicAdd(&A.counters[0], f);
            if (mv) atomicAdd(&A.counters[1], mv);
        }
    }
    if (A.store_state) {
        tile_store<TB, TMA>(gtile, sm, lane);
        if (TMA && lane == 0) bulk_wait_all();
    }
}

// ------------------------------------------------------------------------------------------
// persistent multi-ply pass (spl_rollout): the tile stays in shared memory for `plies` plies of
// uniformly random legal play with Philox reveals and auto reset
// ------------------------------------------------------------------------------------------
struct RolloutParams {
    spl_rollout_args a;
    SplRules rules;
    int n_tiles;
};

// swap_players by ONE seat for the WHOLE lane tile (every lane just moved and passes the turn: the persistent ply loop), as a job of
// the warp: a block of SIZE rows is SIZE x 224 contiguous bytes of the tile (row-major cells, 32 lanes per cell), rolling it by SHIFT
// rows moves 32-bit words - ~40 LDS.32 + 40 STS.32 per lane for two players instead of ~450 byte accesses per lane (spl_rotate walks
// every cell of every lane). Same row permutation as spl_roll_block<SIZE, SHIFT>: new[j] = old[(j + SHIFT) mod SIZE].
template <int SIZE, int SHIFT>
__device__ __forceinline__ void tile_roll_block(int8_t* sm, int row0, int lane) {
    if (SIZE <= 1 || SHIFT % SIZE == 0) return;
    constexpr int W = SIZE * 7 * TL / 4, SW = (SHIFT % SIZE) * 7 * TL / 4, PER = (W + 31) / 32;
    uint32_t* blk = reinterpret_cast<uint32_t*>(sm + (size_t)row0 * 7 * TL);
    uint32_t v[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int d = lane + 32 * i;
        int src = d + SW;
        src = src >= W ? src - W : src;
        v[i] = d < W ? blk[src] : 0u;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int d = lane + 32 * i;
        if (d < W) blk[d] = v[i];
    }
    __syncwarp();
}
template <int N>
static __device__ __noinline__ void lane_rotate_cold(int8_t* lane_base, int k, SplRules r) {   // the rare per-lane path, kept out of the hot loop's code
    TileAcc s{lane_base};
    spl_rotate<N>(s, k, r);
}
template <int N>
__device__ __forceinline__ void tile_rotate1(int8_t* sm, int lane, SplRules r) {
    typedef SplLay<N> L;
    __syncwarp();
    tile_roll_block<N, 1>(sm, L::PGEMS, lane);
    if ((r.flags & SPL_F_REFCOMPAT) || N == 2) tile_roll_block<N * (N + 1), 3 % (N * (N + 1))>(sm, L::PNOBLES, lane);   // :345 (F7a)
    else tile_roll_block<N * (N + 1), N + 1>(sm, L::PNOBLES, lane);
    tile_roll_block<N, 1>(sm, L::PCARDS, lane);
    tile_roll_block<6 * N, 6>(sm, L::PRES, lane);
}

template <int N, bool TMA, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) spl_rollout_kernel(const RolloutParams P) {
    typedef SplLay<N> L;
    constexpr int TB = L::CELLS * TL;
    extern __shared__ __align__(128) int8_t smem[];
    __shared__ uint64_t bars[WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int8_t* sm = smem + (size_t)warp * TB;
    if (TMA) {
        if (lane == 0) mbar_init(&bars[warp], 1);
        __syncwarp();
    }
    const spl_rollout_args& A = P.a;
    uint32_t parity = 0;
    unsigned fin_acc = 0, ply_acc = 0;
    const int warp_stride = gridDim.x * WARPS;
    for (int tile = blockIdx.x * WARPS + warp; tile < P.n_tiles; tile += warp_stride) {
        const int gl = tile * TL + lane;
        const bool active = gl < A.n_lanes;
        int8_t* gtile = A.planes + (size_t)tile * TB;
        tile_load<TB, TMA>(sm, gtile, &bars[warp], parity, lane);
        parity ^= 1;
        TileAcc s{sm + lane};
        const uint32_t game = A.game_base + (uint32_t)gl;
        uint32_t episode = 0;
        int cur = 0;
        if (active) {
            if (A.episodes) episode = A.episodes[gl];
            if (A.players) cur = A.players[gl];
        }
        int first_plies = 0;
        bool have_first = false;
        for (int it = 0; it < A.plies; it++) {   // every lane of the warp runs the loop (padding lanes idle): the rotation and the restart below are warp jobs
            bool restart = false;
            int nxt = 0;
            uint32_t ply = 0;
            if (active) {
                uint32_t m[SPL_MASK_WORDS];
                spl_valid_mask<N>(s, cur, P.rules, m);
                ply = (uint32_t)(uint8_t)s.get(L::BANK, SPL_PTS);
                const int a = spl_pick_random(m, A.seed, game, episode, ply);
                SplChance ch;
                ch.mode = SPL_CHANCE_PHILOX; ch.code = 0; ch.seed = A.seed; ch.game = game; ch.episode = episode; ch.ply = ply;
                nxt = spl_apply_move<N>(s, a, cur, ch);
                if (nxt < 0) nxt = (cur + 1) % N;   // cannot happen for a legal action
                ply_acc++;
            }
            if (A.rotate) {
                // canonical lanes (cur == 0) all pass the turn to seat 1: one rotation of the whole tile by the warp (padding lanes hold zeros);
                // anything else (lanes that came in with another player to move) takes the per-lane path
                if (__all_sync(0xffffffffu, !active || nxt == 1)) tile_rotate1<N>(sm, lane, P.rules);
                else if (active) lane_rotate_cold<N>(sm + lane, nxt, P.rules);
                __syncwarp();
                nxt = 0;
            }
            if (active) {
                cur = nxt;
                float res[N];
                if (spl_game_ended<N>(s, P.rules, res)) {
                    fin_acc++;
                    if (!have_first) {
                        have_first = true;
                        first_plies = (int)ply + 1;
                        if (A.first_result) {
#pragma unroll
                            for (int i = 0; i < N; i++) A.first_result[(size_t)gl * N + i] = res[i];
                        }
                    }
                    episode += 1;
                    cur = 0;
                    restart = true;
                }
            }
            const uint32_t need = __ballot_sync(0xffffffffu, restart);
            if (need) tile_reset_lanes<N>(sm, need, lane, A.seed, game, episode);
        }
        if (active) {
            if (A.episodes) A.episodes[gl] = episode;
            if (A.players) A.players[gl] = (uint8_t)cur;
            if (A.first_plies) A.first_plies[gl] = have_first ? first_plies : 0;
        }
        tile_store<TB, TMA>(gtile, sm, lane);
        tile_store_drain<TMA>(lane);
    }
    if (A.counters) {
        unsigned long long f = warp_sum(fin_acc), mv = warp_sum(ply_acc);
        if (lane == 0) {
            if (f) atomicAdd(&A.counters[0], f);
            if (mv) atomicAdd(&A.counters[1], mv);
        }
    }
    if (TMA && lane == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------
// resets
// ------------------------------------------------------------------------------------------
template <int N, bool TMA, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) spl_reset_kernel(int8_t* planes, int n_lanes, int n_tiles, int explicit_mode,
                                                               uint64_t seed, uint32_t game_base, const uint32_t* episodes,
                                                               const uint8_t* lane_select, const uint8_t* deals, const uint8_t* nobles) {
    typedef SplLay<N> L;
    constexpr int TB = L::CELLS * TL;
    extern __shared__ __align__(128) int8_t smem[];
    __shared__ uint64_t bars[WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * WARPS + warp;
    if (tile >= n_tiles) return;
    int8_t* sm = smem + (size_t)warp * TB;
    if (TMA) {
        if (lane == 0) mbar_init(&bars[warp], 1);
        __syncwarp();
    }
    int8_t* gtile = planes + (size_t)tile * TB;
    tile_load<TB, TMA>(sm, gtile, &bars[warp], 0, lane);   // partial resets keep the other lanes
    const int gl = tile * TL + lane;
    TileAcc s{sm + lane};
    if (gl < n_lanes) {
        if (explicit_mode) {
            spl_init_explicit<N>(s, deals + (size_t)gl * 12, nobles + (size_t)gl * 5);
        } else if (!lane_select || lane_select[gl]) {
            spl_init_philox<N>(s, seed, game_base + (uint32_t)gl, episodes ? episodes[gl] : 0u);
        }
    } else {
        for (int c = 0; c < L::CELLS; c++) sm[c * TL + lane] = 0;   // padding lanes stay all-zero
    }
    tile_store<TB, TMA>(gtile, sm, lane);
    if (TMA && lane == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------
// AoS <-> lane tiles, mask planes -> bool[406], scores
// ------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) spl_pack_kernel(const int8_t* __restrict__ aos, int8_t* __restrict__ planes, int n_lanes) {
    typedef SplLay<N> L;
    constexpr int S = L::CELLS;
    __shared__ int8_t buf[S * TL + 16];
    const int tile = blockIdx.x;
    const int lanes_here = min(TL, n_lanes - tile * TL);
    const int8_t* src = aos + (size_t)tile * TL * S;
    for (int i = threadIdx.x; i < S * TL; i += blockDim.x) buf[i] = i < lanes_here * S ? src[i] : (int8_t)0;
    __syncthreads();
    int8_t* dst = planes + (size_t)tile * S * TL;
    for (int i = threadIdx.x; i < S * TL; i += blockDim.x) {
        const int cell = i >> 5, l = i & 31;
        dst[i] = buf[l * S + cell];
    }
}
template <int N>
__global__ void __launch_bounds__(128) spl_unpack_kernel(const int8_t* __restrict__ planes, int8_t* __restrict__ aos, int n_lanes) {
    typedef SplLay<N> L;
    constexpr int S = L::CELLS;
    __shared__ int8_t buf[S * (TL + 1) + 16];
    const int tile = blockIdx.x;
    const int lanes_here = min(TL, n_lanes - tile * TL);
    const int8_t* src = planes + (size_t)tile * S * TL;
    for (int i = threadIdx.x; i < S * TL; i += blockDim.x) {
        const int cell = i >> 5, l = i & 31;
        buf[cell * (TL + 1) + l] = src[i];
    }
    __syncthreads();
    int8_t* dst = aos + (size_t)tile * TL * S;
    for (int i = threadIdx.x; i < lanes_here * S; i += blockDim.x) {
        const int l = i / S, cell = i - l * S;
        dst[i] = buf[cell * (TL + 1) + l];
    }
}

__global__ void spl_mask_unpack_kernel(const uint32_t* __restrict__ mask, uint8_t* __restrict__ valids, int n_lanes, size_t lpad) {
    const size_t total = (size_t)n_lanes * SPL_ACTIONS;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t l = i / SPL_ACTIONS;
        const int a = (int)(i - l * SPL_ACTIONS);
        valids[i] = (uint8_t)((mask[(size_t)(a >> 5) * lpad + l] >> (a & 31)) & 1u);
    }
}

template <int N>
__global__ void spl_scores_kernel(const int8_t* __restrict__ planes, int n_lanes, SplRules rules, int32_t* scores, int32_t* rounds) {
    typedef SplLay<N> L;
    const int gl = blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= n_lanes) return;
    GTileAcc s{planes + (size_t)(gl >> 5) * L::CELLS * TL + (gl & 31)};
    if (scores) {
#pragma unroll
        for (int p = 0; p < N; p++) scores[(size_t)gl * N + p] = spl_score<N>(s, p, rules);
    }
    if (rounds) rounds[gl] = (int)(uint8_t)s.get(L::BANK, SPL_PTS);
}

// ------------------------------------------------------------------------------------------
// get_symmetries (:349-395) on AoS input; one CTA per state
// ------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) spl_sym_kernel(const int8_t* __restrict__ aos, const float* __restrict__ pi,
                                                      const uint8_t* __restrict__ valids, int8_t* out_states, float* out_pi,
                                                      uint8_t* out_valids, int32_t* out_count) {
    typedef SplLay<N> L;
    constexpr int S = L::CELLS, MAXV = SPL_MAX_SYMMETRIES;
    __shared__ int8_t st[S];
    __shared__ int8_t kind[MAXV], arg0[MAXV], perm[MAXV][4];
    __shared__ int count;
    const size_t g = blockIdx.x;
    for (int i = threadIdx.x; i < S; i += blockDim.x) st[i] = aos[g * S + i];
    __syncthreads();
    if (threadIdx.x == 0) {
        int k = 0;
        kind[k++] = 0;                                                   // identity :367
        for (int tier = 0; tier < 3; tier++)
            for (int j = 0; j < 3; j++) {                                // visible-card permutations :369-376
                kind[k] = 1; arg0[k] = (int8_t)tier;
                for (int q = 0; q < 4; q++) perm[k][q] = SPL_CARD_SYM[j][q];
                k++;
            }
        for (int p = 0; p < N; p++) {                                    // reserved-card permutations :379-393
            int nb = 3;
            for (int c = 2; c >= 0; c--) {
                int t = 0;
                for (int q = 0; q < 5; q++) t += st[(L::PRES + 6 * p + 2 * c) * 7 + q];
                if (t == 0) nb = c;
            }
            for (int j = 0; j < 2; j++) {
                if (SPL_RES_SYM[nb][j][0] < 0) continue;
                kind[k] = 2; arg0[k] = (int8_t)p;
                for (int q = 0; q < 3; q++) perm[k][q] = SPL_RES_SYM[nb][j][q];
                perm[k][3] = 3;
                k++;
            }
        }
        count = k;
        out_count[g] = k;
    }
    __syncthreads();
    for (int v = 0; v < count; v++) {
        const int kd = kind[v], a0 = arg0[v];
        int8_t* os = out_states + (g * MAXV + v) * S;
        float* op = out_pi + (g * MAXV + v) * SPL_ACTIONS;
        uint8_t* ov = out_valids + (g * MAXV + v) * SPL_ACTIONS;
        const int row0 = kd == 1 ? L::CARDS + 8 * a0 : L::PRES + 6 * a0;   // first permuted row
        const int nslots = kd == 1 ? 4 : 3;
        for (int i = threadIdx.x; i < S; i += blockDim.x) {
            int src = i;
            if (kd) {
                const int row = i / 7, rel = row - row0;
                if (rel >= 0 && rel < 2 * nslots) src = (row0 + 2 * perm[v][rel >> 1] + (rel & 1)) * 7 + (i - row * 7);
            }
            os[i] = st[src];
        }
        for (int a = threadIdx.x; a < SPL_ACTIONS; a += blockDim.x) {
            int src = a;
            if (kd == 1) {
                if (a >= 4 * a0 && a < 4 * a0 + 4) src = 4 * a0 + perm[v][a - 4 * a0];
                else if (a >= 12 + 4 * a0 && a < 16 + 4 * a0) src = 12 + 4 * a0 + perm[v][a - 12 - 4 * a0];
            } else if (kd == 2 && a0 == 0) {
                if (a >= 27 && a < 30) src = 27 + perm[v][a - 27];
            }
            op[a] = pi[g * SPL_ACTIONS + src];
            ov[a] = valids[g * SPL_ACTIONS + src];
        }
    }
}

// ==========================================================================================
// host side of the C ABI
// ==========================================================================================
template <int N> struct Cfg {   // small CTAs (2 warps = 2 tiles) so that shared memory packs tightly: 9 / 7 / 5 CTAs per SM
    static constexpr int WARPS = 2;
    static constexpr int SMEM = WARPS * SplLay<N>::CELLS * TL;
    static constexpr int MINB = (227 * 1024) / (SMEM + 1024);
};

template <typename K>
static cudaError_t set_smem(K kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

extern "C" {

int spl_abi_version(void) { return SPL_ABI_VERSION; }
const char* spl_last_error(void) { return g_err; }

int spl_state_rows(int n) { return 32 + 10 * n + n * n; }
int spl_state_bytes(int n) { return 7 * spl_state_rows(n); }
int spl_lanes_padded(int n_lanes) { return (n_lanes + TL - 1) / TL * TL; }
size_t spl_planes_bytes(int n, int n_lanes) { return (size_t)spl_lanes_padded(n_lanes) * (size_t)spl_state_bytes(n); }
size_t spl_mask_planes_bytes(int n_lanes) { return (size_t)spl_lanes_padded(n_lanes) * SPL_MASK_WORDS * 4; }

int spl_ctx_create(int n_players, int token_limit, uint32_t rule_flags, int device, spl_ctx** out) {
    if (!out || n_players < 2 || n_players > 4 || token_limit < 1 || token_limit > 10) return fail(SPL_E_ARG, "spl_ctx_create: bad argument");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(SPL_E_NOGPU, "spl_ctx_create: no CUDA device (there is no CPU fallback)", e);
    if (device < 0 || device >= count) return fail(SPL_E_ARG, "spl_ctx_create: bad device index");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(SPL_E_NOGPU, "spl_ctx_create: kernels are built for sm_100a only");
    spl_ctx* c = new spl_ctx;
    c->n = n_players; c->rules.limit = token_limit; c->rules.flags = rule_flags; c->device = device; c->use_tma = 1;
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return SPL_OK;
}
int spl_ctx_set_rules(spl_ctx* c, int token_limit, uint32_t rule_flags) {
    if (!c || token_limit < 1 || token_limit > 10) return fail(SPL_E_ARG, "spl_ctx_set_rules: bad argument");
    c->rules.limit = token_limit; c->rules.flags = rule_flags;
    return SPL_OK;
}
int spl_ctx_set_tma(spl_ctx* c, int enabled) {
    if (!c) return fail(SPL_E_ARG, "spl_ctx_set_tma: null context");
    c->use_tma = enabled ? 1 : 0;
    return SPL_OK;
}
void spl_ctx_destroy(spl_ctx* c) { delete c; }

#define ENTER(c)                                                   \
    if (!(c)) return fail(SPL_E_ARG, "null context");              \
    CU(cudaSetDevice((c)->device));                                \
    cudaStream_t st = (cudaStream_t)stream;

int spl_pack(spl_ctx* c, const int8_t* aos, int8_t* planes, int n_lanes, void* stream) {
    ENTER(c);
    if (!aos || !planes || n_lanes <= 0) return fail(SPL_E_ARG, "spl_pack: bad argument");
    const int tiles = spl_lanes_padded(n_lanes) / TL;
    DISPATCH_N(c->n, spl_pack_kernel<N><<<tiles, 128, 0, st>>>(aos, planes, n_lanes));
    CU(cudaGetLastError());
    return SPL_OK;
}
int spl_unpack(spl_ctx* c, const int8_t* planes, int8_t* aos, int n_lanes, void* stream) {
    ENTER(c);
    if (!aos || !planes || n_lanes <= 0) return fail(SPL_E_ARG, "spl_unpack: bad argument");
    const int tiles = spl_lanes_padded(n_lanes) / TL;
    DISPATCH_N(c->n, spl_unpack_kernel<N><<<tiles, 128, 0, st>>>(planes, aos, n_lanes));
    CU(cudaGetLastError());
    return SPL_OK;
}
int spl_mask_unpack(spl_ctx* c, const uint32_t* mask_planes, uint8_t* valids, int n_lanes, void* stream) {
    ENTER(c);
    if (!mask_planes || !valids || n_lanes <= 0) return fail(SPL_E_ARG, "spl_mask_unpack: bad argument");
    const size_t total = (size_t)n_lanes * SPL_ACTIONS;
    const int blocks = (int)((total + 255) / 256 < (size_t)(c->sm_count * 16) ? (total + 255) / 256 : (size_t)(c->sm_count * 16));
    spl_mask_unpack_kernel<<<blocks, 256, 0, st>>>(mask_planes, valids, n_lanes, (size_t)spl_lanes_padded(n_lanes));
    CU(cudaGetLastError());
    return SPL_OK;
}

}   // extern "C"
template <int N, bool TMA>
static int launch_reset(spl_ctx* c, int8_t* planes, int n_lanes, int explicit_mode, uint64_t seed, uint32_t game_base,
                        const uint32_t* episodes, const uint8_t* lane_select, const uint8_t* deals, const uint8_t* nobles, cudaStream_t st) {
    constexpr int W = Cfg<N>::WARPS;
    const int tiles = spl_lanes_padded(n_lanes) / TL;
    auto k = spl_reset_kernel<N, TMA, W>;
    CU(set_smem(k, Cfg<N>::SMEM));
    k<<<(tiles + W - 1) / W, W * 32, Cfg<N>::SMEM, st>>>(planes, n_lanes, tiles, explicit_mode, seed, game_base, episodes, lane_select, deals, nobles);
    CU(cudaGetLastError());
    (void)c;
    return SPL_OK;
}
extern "C" {

int spl_reset_philox(spl_ctx* c, int8_t* planes, int n_lanes, uint64_t seed, uint32_t game_base, const uint32_t* episodes,
                     const uint8_t* lane_select, void* stream) {
    ENTER(c);
    if (!planes || n_lanes <= 0) return fail(SPL_E_ARG, "spl_reset_philox: bad argument");
    int rc = 0;
    DISPATCH_N(c->n, rc = c->use_tma ? launch_reset<N, true>(c, planes, n_lanes, 0, seed, game_base, episodes, lane_select, nullptr, nullptr, st)
                                     : launch_reset<N, false>(c, planes, n_lanes, 0, seed, game_base, episodes, lane_select, nullptr, nullptr, st));
    return rc;
}
int spl_reset_explicit(spl_ctx* c, int8_t* planes, int n_lanes, const uint8_t* deals, const uint8_t* nobles, void* stream) {
    ENTER(c);
    if (!planes || !deals || !nobles || n_lanes <= 0) return fail(SPL_E_ARG, "spl_reset_explicit: bad argument");
    int rc = 0;
    DISPATCH_N(c->n, rc = c->use_tma ? launch_reset<N, true>(c, planes, n_lanes, 1, 0, 0, nullptr, nullptr, deals, nobles, st)
                                     : launch_reset<N, false>(c, planes, n_lanes, 1, 0, 0, nullptr, nullptr, deals, nobles, st));
    return rc;
}

}   // extern "C"
template <int N, bool TMA>
static int launch_step(spl_ctx* c, const spl_step_args* a, cudaStream_t st) {
    constexpr int W = Cfg<N>::WARPS;
    StepParams P;
    P.a = *a; P.rules = c->rules; P.n_tiles = spl_lanes_padded(a->n_lanes) / TL;
    auto k = spl_step_kernel<N, TMA, W, Cfg<N>::MINB>;
    CU(set_smem(k, Cfg<N>::SMEM));
    k<<<(P.n_tiles + W - 1) / W, W * 32, Cfg<N>::SMEM, st>>>(P);
    CU(cudaGetLastError());
    return SPL_OK;
}
extern "C" {
int spl_step(spl_ctx* c, const spl_step_args* a, void* stream) {
    ENTER(c);
    if (!a || !a->planes || a->n_lanes <= 0) return fail(SPL_E_ARG, "spl_step: bad argument");
    if (a->chance_mode < 0 || a->chance_mode > 2) return fail(SPL_E_ARG, "spl_step: bad chance_mode");
    if (a->chance_mode == SPL_CHANCE_REPLAY && !a->reveals) return fail(SPL_E_ARG, "spl_step: replay needs reveals");
    if (a->auto_reset && !a->episodes) return fail(SPL_E_ARG, "spl_step: auto_reset needs episodes");
    if (!a->players && (a->player < 0 || a->player >= c->n)) return fail(SPL_E_ARG, "spl_step: bad player");
    int rc = 0;
    DISPATCH_N(c->n, rc = c->use_tma ? launch_step<N, true>(c, a, st) : launch_step<N, false>(c, a, st));
    return rc;
}

}   // extern "C"
template <int N, bool TMA>
static int launch_rollout(spl_ctx* c, const spl_rollout_args* a, cudaStream_t st) {
    constexpr int W = Cfg<N>::WARPS;
    RolloutParams P;
    P.a = *a; P.rules = c->rules; P.n_tiles = spl_lanes_padded(a->n_lanes) / TL;
    auto k = spl_rollout_kernel<N, TMA, W, Cfg<N>::MINB>;
    CU(set_smem(k, Cfg<N>::SMEM));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, W * 32, Cfg<N>::SMEM));
    if (per_sm < 1) per_sm = 1;
    int grid = c->sm_count * per_sm;                 // persistent: every SM full, warps stride over tiles
    const int need = (P.n_tiles + W - 1) / W;
    if (grid > need) grid = need;
    k<<<grid, W * 32, Cfg<N>::SMEM, st>>>(P);
    CU(cudaGetLastError());
    return SPL_OK;
}
extern "C" {
int spl_rollout(spl_ctx* c, const spl_rollout_args* a, void* stream) {
    ENTER(c);
    if (!a || !a->planes || a->n_lanes <= 0 || a->plies < 0) return fail(SPL_E_ARG, "spl_rollout: bad argument");
    if (!a->rotate && !a->players) return fail(SPL_E_ARG, "spl_rollout: rotate=0 needs the players array");
    int rc = 0;
    DISPATCH_N(c->n, rc = c->use_tma ? launch_rollout<N, true>(c, a, st) : launch_rollout<N, false>(c, a, st));
    return rc;
}

int spl_scores(spl_ctx* c, const int8_t* planes, int n_lanes, int32_t* scores, int32_t* rounds, void* stream) {
    ENTER(c);
    if (!planes || n_lanes <= 0) return fail(SPL_E_ARG, "spl_scores: bad argument");
    DISPATCH_N(c->n, spl_scores_kernel<N><<<(n_lanes + 127) / 128, 128, 0, st>>>(planes, n_lanes, c->rules, scores, rounds));
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_symmetries(spl_ctx* c, const int8_t* aos, const float* pi, const uint8_t* valids, int n_lanes, int8_t* out_states,
                   float* out_pi, uint8_t* out_valids, int32_t* out_count, void* stream) {
    ENTER(c);
    if (!aos || !pi || !valids || !out_states || !out_pi || !out_valids || !out_count || n_lanes <= 0)
        return fail(SPL_E_ARG, "spl_symmetries: bad argument");
    DISPATCH_N(c->n, spl_sym_kernel<N><<<n_lanes, 128, 0, st>>>(aos, pi, valids, out_states, out_pi, out_valids, out_count));
    CU(cudaGetLastError());
    return SPL_OK;
}

}   // extern "C"
