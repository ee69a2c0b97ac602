// libsplendor_b200.so - MCTS tree arena kernels for sm_100a and their C ABI (include/splendor_b200.h).
// The per-tree logic lives in spl_mcts.cuh; this file is launch plumbing. A selection wave is
//   mcts_descend_kernel / mcts_expand_descend_kernel   one warp per tree: (expansion of the previous leaf +) PUCT descent + path;
//                        with RULES also the rules step of the tree's pending edge (lane 0 of the warp)
//   mcts_rules_kernel    the rules step as its own launch: one LANE per tree, a few trees per warp, states staged by cp.async:
//                        make_move + swap_players + getGameEnded + getValidMoves of the child of every pending edge
//   mcts_attach_kernel   one warp per tree: hash, dictionary lookup / insertion, edge allocation, leaf hand-over
// spl_mcts_wave_nnet chains them with the fused evaluator (spl_nnet.cu): descent -> network, attach on a side stream.
// A tree whose new edge led into a node it already holds (transposition) or into a terminal node carries on in the next
// wave; `rounds` > 1 repeats the three kernels inside one wave instead (measured: not worth the extra straggler-bound launches).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "spl_internal.h"
#include "spl_mcts.cuh"

#ifndef MW
#define MW 1                 // warps (trees) per CTA: one, so that a CTA slot is free again as soon as ITS tree is done (measured +3 % over 4)
#endif
#ifndef DESC_MINB
#define DESC_MINB 8          // resident 4-warp groups per SM the descent kernel is compiled for (8 -> 64 registers)
#endif
// simulations that end in a terminal node a descend call backs up before it yields: one in lock-step calls (max_levels = 0: the wave lasts as long
// as its slowest tree), four when descents yield anyway (asynchronous moves; measured +1.5 % there, -30 % on the lock-step call)
#define DESC_MAX_TERMINAL (max_levels < (1 << 20) ? 4 : 1)   // (the host turns max_levels = 0 into 1 << 20)
#ifndef ATT_MINB
#define ATT_MINB 8
#endif
#define MINB(b) ((b) * 4 / MW)   // launch bounds are stated per four warps, whatever MW is
#define MSP 640              // per-warp state scratch (>= MctsLay<4>::SP = 624)

struct spl_mcts {
    spl_ctx* ctx;
    MctsArena A;
    MctsSearchParams P;
    int gc_reachable, rounds, max_levels;
    const uint32_t* episodes;    // the lanes' episode counters (device, may be NULL): part of the Philox key of the on-device Dirichlet sampler
    int rules_tpw;               // trees per warp of the rules kernel
    int pdl;                     // spl_mcts_wave_nnet: descent and network as programmatic dependent launches on one stream
    int fuse_rules;              // spl_mcts_wave_nnet: rules step inside the descent kernel (no launch boundary) instead of mcts_rules_kernel
    cudaStream_t side;           // spl_mcts_wave_nnet: the network runs here, next to the attach kernel
    cudaEvent_t ev_fork, ev_join;
};

// diagnostics: per-tree time stamps when the arena carries a profile buffer (spl_mcts_debug_profile), lane 0 of the tree's warp
__device__ __forceinline__ long long prof_globaltimer() {
    unsigned long long x;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(x));
    return (long long)x;
}
#define PROF_STAMP(A, t, i, val) do { if ((A).prof && (threadIdx.x & 31) == 0) (A).prof[(size_t)(t) * 16 + (i)] = (val); } while (0)

struct WarpScratch {
    int8_t* st;
    uint32_t* words;
    double* dwords;
    uint8_t* cs;     // 128 bytes: a compact node state
};

__device__ __forceinline__ WarpScratch warp_scratch(int warp) {
    __shared__ __align__(16) int8_t s_state[MW][MSP];
    __shared__ uint32_t s_words[MW][24];
    __shared__ double s_dwords[MW][4];
    __shared__ __align__(16) uint8_t s_cs[MW][128];
    WarpScratch s;
    s.st = s_state[warp]; s.words = s_words[warp]; s.dwords = s_dwords[warp]; s.cs = s_cs[warp];
    return s;
}

template <int N>
__global__ void __launch_bounds__(MW * 32) mcts_begin_kernel(const __grid_constant__ MctsArena A, const __grid_constant__ MctsSearchParams P, const int8_t* roots, const int32_t* sims,
                                                             const uint8_t* move_flags, const uint8_t* tree_select, const double* dir,
                                                             const uint32_t* episodes, int gc_reachable) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    WarpScratch sc = warp_scratch(warp);
    if (t >= A.n_trees) return;
    if (tree_select && !tree_select[t]) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_begin_tree<N>(w, A, t, P, roots + (size_t)t * MctsLay<N>::S, sims[t], move_flags ? (uint32_t)move_flags[t] : 0u, gc_reachable,
                       dir ? dir + (size_t)t * SPL_ACTIONS : nullptr, episodes ? episodes[t] : 0u, sc.st, sc.cs, sc.words, sc.dwords);
}

template <int N, bool VL>
__global__ void __launch_bounds__(MW * 32, MINB(7)) mcts_descend_kernel(const __grid_constant__ MctsArena A, const __grid_constant__ MctsSearchParams P, int max_levels, int8_t* leaf_states,
                                                                  uint8_t* leaf_valids) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    for (int s = 0; s < (VL ? A.n_slots : 1); s++) {
        const size_t r = (size_t)mcts_row(A, t, s);
        mcts_descend_tree<N, VL>(w, A, t, s, P, DESC_MAX_TERMINAL, max_levels, leaf_states + r * MctsLay<N>::S, leaf_valids + r * SPL_ACTIONS);
        __syncwarp();
    }
}

#define RW 2   // warps per CTA of the rules kernel

// One LANE per tree, TPW trees per warp (the other lanes only help with the copies). The kernel is bound by the latency of a
// single warp walking through the rules code, so everything around that walk is kept short: the parent states (the reference's
// own byte order, sp bytes each) come in with one batch of 16-byte cp.async per warp - all of them in flight at once - and
// stay in that order in shared memory (row stride chosen so that the TPW lanes hit different banks), the rules code reads and
// writes them through the AoS accessor, and the children go out to the staging rows as straight 16-byte copies.
template <int N> struct RulesSmem {
    static constexpr int SP = MctsLay<N>::SP;
    static constexpr int STRIDE = (SP / 4) % 32 == 0 ? SP + 16 : SP;   // 432 / 528 / 624 bytes: lane strides of 12 / 4 / 28 banks
};

template <int N, int TPW>
__global__ void __launch_bounds__(RW * 32) mcts_rules_kernel(MctsArena A, SplRules rules) {
    typedef MctsLay<N> ML;
    constexpr int STRIDE = RulesSmem<N>::STRIDE, CH = ML::SP / 16;
    extern __shared__ __align__(16) int8_t tile_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_rows = A.n_trees * A.n_slots;            // a row = (tree, in-flight slot)
    const int t0 = (blockIdx.x * RW + warp) * TPW, t = t0 + lane;
    int8_t* wsm = tile_smem + (size_t)warp * (TPW * (STRIDE + 128) + 64);
    uint8_t* csm = reinterpret_cast<uint8_t*>(wsm + TPW * STRIDE);      // the parents' compact states, 128 bytes each
    uint32_t* psm = reinterpret_cast<uint32_t*>(csm + TPW * 128);       // their record offsets
    bool pending = false;
    uint32_t parent = 0u;
    int action = 0;
    if (lane < TPW && t < n_rows) {
        const MctsSlot* T = A.trees[t / A.n_slots].slot + t % A.n_slots;
        const int pe = T->pend_edge;
        if (pe >= 0 && T->leaf == 0u) {
            pending = true;
            parent = T->pend_parent;
            action = (int)mcts_edges(A, parent, pe >> 16).ca[pe & 0xFFFF].action;
        }
    }
    const uint32_t pmask = __ballot_sync(0xffffffffu, pending);
    if (pmask == 0u) return;
    const bool prof_ok = A.n_slots == 1 && t0 < A.n_trees;
    if (prof_ok) { PROF_STAMP(A, t0, 6, prof_globaltimer()); PROF_STAMP(A, t0, 7, clock64()); }
    {   // the parents' states: records hold the compact form, the rules code reads the reference's layout. All compact states of the
        // warp's trees come in with ONE batch of 16-byte loads (one memory round trip), then the whole warp decodes tree after tree.
        const MctsWarp w{lane};
        constexpr int C16 = MctsCLay<N>::CP / 16;
        if (lane < TPW) psm[lane] = parent;
        __syncwarp();
        for (int i = lane; i < TPW * C16; i += 32) {
            const int j = i / C16, c = i - j * C16;
            if ((pmask >> j) & 1u) reinterpret_cast<uint4*>(csm + 128 * j)[c] = reinterpret_cast<const uint4*>(mcts_cstate(A, psm[j]))[c];
        }
        uint4 z; z.x = z.y = z.z = z.w = 0u;
        for (int i = lane; i < TPW * CH; i += 32) {
            const int j = i / CH, c = i - j * CH;
            reinterpret_cast<uint4*>(wsm + j * STRIDE)[c] = z;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < TPW; j++)
            if ((pmask >> j) & 1u) mcts_decode<N>(w, csm + 128 * j, wsm + j * STRIDE, true);
        __syncwarp();
    }
    if (prof_ok) PROF_STAMP(A, t0, 8, clock64());
    bool ended = false;
    float es[N];
    uint32_t m[SPL_MASK_WORDS];
    if (pending) {
        AosAcc s{wsm + lane * STRIDE};
        ended = mcts_rules_core<N>(s, action, rules, es, m);
    }
    __syncwarp();
    if (prof_ok) PROF_STAMP(A, t0, 9, clock64());
#pragma unroll
    for (int j = 0; j < TPW; j++) {
        if ((pmask >> j) & 1u) {
            uint4* dst = reinterpret_cast<uint4*>(A.stage_state + (size_t)(t0 + j) * A.sp);
            const uint4* src = reinterpret_cast<const uint4*>(wsm + j * STRIDE);
            for (int i = lane; i < CH; i += 32) dst[i] = src[i];
        }
    }
    if (pending) {
#pragma unroll
        for (int i = 0; i < SPL_MASK_WORDS; i++) A.stage_mask[(size_t)i * n_rows + t] = m[i];
#pragma unroll
        for (int i = 0; i < 4; i++) A.stage_es[(size_t)t * 4 + i] = i < N ? es[i] : 0.f;
        A.stage_ended[t] = ended ? 1 : 0;
        A.leaf_src[t] = 1;   // if this child needs the network, its input row is the staging row (spl_mcts_wave_nnet)
    }
    if (prof_ok) { PROF_STAMP(A, t0, 10, clock64()); PROF_STAMP(A, t0, 11, prof_globaltimer()); }
}

template <int N>
static cudaError_t launch_rules(const MctsArena& A, const SplRules& rules, int tpw, cudaStream_t st) {
    const int warps = (A.n_trees * A.n_slots + tpw - 1) / tpw, grid = (warps + RW - 1) / RW;
    const int smem = RW * (tpw * (RulesSmem<N>::STRIDE + 128) + 64);
    switch (tpw) {
        case 16: mcts_rules_kernel<N, 16><<<grid, RW * 32, smem, st>>>(A, rules); break;
        case 8: mcts_rules_kernel<N, 8><<<grid, RW * 32, smem, st>>>(A, rules); break;
        case 4: mcts_rules_kernel<N, 4><<<grid, RW * 32, smem, st>>>(A, rules); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int N, bool VL>
__global__ void __launch_bounds__(MW * 32, MINB(ATT_MINB)) mcts_attach_kernel(const __grid_constant__ MctsArena A, const __grid_constant__ MctsSearchParams P, int8_t* leaf_states, uint8_t* leaf_valids,
                                                                 uint8_t* leaf_flags, int32_t* counters, bool emit_rows) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    __shared__ __align__(16) uint8_t s_cs[MW][128];
    __shared__ __align__(16) int8_t s_aos[MW][MSP];
    MctsWarp w{(int)(threadIdx.x & 31)};
    PROF_STAMP(A, t, 12, prof_globaltimer()); PROF_STAMP(A, t, 13, clock64());
    const int n_rows = A.n_trees * A.n_slots;
    int leaves = 0;
    bool waiting = false;
    for (int s = 0; s < (VL ? A.n_slots : 1); s++) {   // the slots of a tree one after the other: node creation stays deterministic
        const size_t r = (size_t)mcts_row(A, t, s);
        if (A.trees[t].slot[s].pend_edge >= 0) {   // the child's bytes: staging row -> shared memory (the encoder reads them cell by cell)
            const uint4* src = reinterpret_cast<const uint4*>(A.stage_state + r * A.sp);
            for (int i = w.lane; i < A.sp / 16; i += 32) reinterpret_cast<uint4*>(s_aos[warp])[i] = src[i];
        }
        __syncwarp();
        const int leaf = mcts_attach_tree<N, VL>(w, A, t, s, P, s_aos[warp], s_cs[warp], A.stage_ended[r] != 0, A.stage_es + r * 4,
                                                 A.stage_mask + r, n_rows, leaf_states + r * MctsLay<N>::S, leaf_valids + r * SPL_ACTIONS, emit_rows);
        __syncwarp();
        if (w.lane == 0) leaf_flags[r] = (uint8_t)leaf;
        leaves += leaf;
        waiting |= A.trees[t].slot[s].pend_edge >= 0 || A.trees[t].slot[s].cur != 0u;
    }
    PROF_STAMP(A, t, 14, clock64()); PROF_STAMP(A, t, 15, prof_globaltimer());
    if (w.lane == 0 && counters) {   // last round of the wave
        const MctsTree* T = A.trees + t;
        if (leaves) atomicAdd(&counters[0], leaves);
        if (T->status == 0u && (leaves || waiting || T->sims_done < T->sims_target)) atomicAdd(&counters[1], 1);
    }
}

template <int N, bool VL>
__global__ void __launch_bounds__(MW * 32, MINB(8)) mcts_expand_kernel(const __grid_constant__ MctsArena A, const __grid_constant__ MctsSearchParams P, const float* pi, const float* v, const double* dir) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    WarpScratch sc = warp_scratch(warp);
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    for (int s = 0; s < (VL ? A.n_slots : 1); s++) {
        const size_t r = (size_t)mcts_row(A, t, s);
        mcts_expand_tree<N, VL>(w, A, t, s, P, pi + r * SPL_ACTIONS, v + r * N, dir ? dir + (size_t)t * SPL_ACTIONS : nullptr, sc.dwords);
        __syncwarp();
    }
}

// rotation of a contiguous block of BYTES bytes by `shift` bytes, all lanes of the warp: new[i] = old[(i + shift) mod BYTES]
template <int BYTES>
__device__ __forceinline__ void coop_roll_rows(int8_t* base, int shift, int lane) {
    int8_t v[(BYTES + 31) / 32];
#pragma unroll
    for (int q = 0; q < (BYTES + 31) / 32; q++) {
        const int i = lane + 32 * q;
        int j = i + shift;
        j = j >= BYTES ? j - BYTES : j;
        v[q] = i < BYTES ? base[j] : (int8_t)0;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < (BYTES + 31) / 32; q++) {
        const int i = lane + 32 * q;
        if (i < BYTES) base[i] = v[q];
    }
    __syncwarp();
}

// the rules step of ONE tree by its own warp (lane 0 walks the rules code, the warp does the copies): what mcts_rules_kernel does
// for a few trees per warp, here without a kernel boundary between the descent and it. wsm: sp bytes of shared memory, 16-aligned.
template <int N>
__device__ __forceinline__ void rules_for_own_tree(const MctsArena& A, int t, const SplRules& rules, int8_t* wsm, int lane) {
    constexpr int CH = MctsLay<N>::SP / 16;
    const MctsSlot* T = A.trees[t].slot;     // (the rules step inside the descent kernel is a one-leaf-per-tree feature: slot 0)
    const int pe = T->pend_edge;
    if (pe < 0 || T->leaf != 0u) return;
    const int action = (int)mcts_edges(A, T->pend_parent, pe >> 16).ca[pe & 0xFFFF].action;
    {
        const MctsWarp w{lane};
        uint4 z; z.x = z.y = z.z = z.w = 0u;
        if (lane < CH) reinterpret_cast<uint4*>(wsm)[lane] = z;
        if (lane + 32 < CH) reinterpret_cast<uint4*>(wsm)[lane + 32] = z;
        __syncwarp();
        mcts_decode<N>(w, mcts_cstate(A, T->pend_parent), wsm, true);
        __syncwarp();
    }
    PROF_STAMP(A, t, 8, clock64());
    // make_move on lane 0; swap_players (a byte rotation of the four per-player blocks) and the 15 candidate cards of valid_moves
    // spread over the lanes; the rest of valid_moves and getGameEnded on lane 0 again - the same arithmetic as mcts_rules_core
    typedef SplLay<N> L;
    int nxt = 0;
    if (lane == 0) {
        AosAcc s{wsm};
        SplChance ch;
        ch.mode = 0; ch.code = 0; ch.seed = 0; ch.game = 0; ch.episode = 0; ch.ply = 0;
        nxt = spl_apply_move<N>(s, action, 0, ch);
    }
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
    __syncwarp();
    for (int it = 0; it < nxt; it++) {          // spl_rotate: new row j of a block = old row (j + shift) mod size, for four blocks
        const int nob_shift = ((rules.flags & SPL_F_REFCOMPAT) || N == 2) ? 3 % (N * (N + 1)) : N + 1;   // :345 (F7a)
        coop_roll_rows<N * 7>(wsm + 7 * L::PGEMS, 7, lane);
        coop_roll_rows<N * (N + 1) * 7>(wsm + 7 * L::PNOBLES, 7 * nob_shift, lane);
        coop_roll_rows<N * 7>(wsm + 7 * L::PCARDS, 7, lane);
        coop_roll_rows<6 * N * 7>(wsm + 7 * L::PRES, 42, lane);
    }
    PROF_STAMP(A, t, 10, clock64());
    uint32_t pre[2] = {0u, 0u};
    if (lane < 15) {
        AosAcc s{wsm};
        int have[5];
#pragma unroll
        for (int c = 0; c < 5; c++) have[c] = s.get(L::PGEMS, c) + s.get(L::PCARDS, c);
        spl_card_bits<N>(s, 0, lane, have, s.get(L::PGEMS, 5), pre[0], pre[1]);
    }
    pre[0] = __reduce_or_sync(0xffffffffu, pre[0]);
    pre[1] = __reduce_or_sync(0xffffffffu, pre[1]);
    PROF_STAMP(A, t, 11, clock64());
    if (lane == 0) {
        AosAcc s{wsm};
        float es[N];
        uint32_t m[SPL_MASK_WORDS];
        const bool ended = spl_game_ended<N>(s, rules, es);
        if (!ended) spl_valid_mask<N>(s, 0, rules, m, pre);
        else
            for (int i = 0; i < SPL_MASK_WORDS; i++) m[i] = 0u;
#pragma unroll
        for (int i = 0; i < SPL_MASK_WORDS; i++) A.stage_mask[(size_t)i * A.n_trees + t] = m[i];
#pragma unroll
        for (int i = 0; i < 4; i++) A.stage_es[(size_t)t * 4 + i] = i < N ? es[i] : 0.f;
        A.stage_ended[t] = ended ? 1 : 0;
        A.leaf_src[t] = 1;
    }
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(A.stage_state + (size_t)t * A.sp);
    for (int i = lane; i < CH; i += 32) dst[i] = reinterpret_cast<const uint4*>(wsm)[i];
}

// expansion of the previous wave's leaf and the next descent of the same tree in one launch (both are warp-per-tree);
// RULES: followed by the rules step of the tree's pending edge (spl_mcts_wave_nnet; otherwise mcts_rules_kernel does it)
template <int N, bool RULES, bool VL>
__global__ void __launch_bounds__(MW * 32, MINB(DESC_MINB)) mcts_expand_descend_kernel(const __grid_constant__ MctsArena A, const __grid_constant__ MctsSearchParams P, const float* pi, const float* v, const double* dir,
                                                                         int max_levels, int8_t* leaf_states, uint8_t* leaf_valids) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    WarpScratch sc = warp_scratch(warp);
    // programmatic dependent launch (spl_mcts_wave_nnet): this grid may become resident while the network of the previous wave is
    // still running; nothing is read before that grid has completed and its pi / v rows are visible. A no-op otherwise.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // once every block of this grid is running, the next grid may move in behind it
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    if (VL) {   // several simulations per tree and wave: all expansions / backups first, then the descents one after the other
        for (int s = 0; s < A.n_slots; s++) {
            const size_t r = (size_t)mcts_row(A, t, s);
            mcts_expand_tree<N, true>(w, A, t, s, P, pi + r * SPL_ACTIONS, v + r * N, dir ? dir + (size_t)t * SPL_ACTIONS : nullptr, sc.dwords);
            __syncwarp();
        }
        for (int s = 0; s < A.n_slots; s++) {
            const size_t r = (size_t)mcts_row(A, t, s);
            mcts_descend_tree<N, true>(w, A, t, s, P, DESC_MAX_TERMINAL, max_levels, leaf_states + r * MctsLay<N>::S, leaf_valids + r * SPL_ACTIONS);
            __syncwarp();
        }
        return;
    }
    PROF_STAMP(A, t, 0, prof_globaltimer()); PROF_STAMP(A, t, 1, clock64());
    mcts_expand_tree<N, false>(w, A, t, 0, P, pi + (size_t)t * SPL_ACTIONS, v + (size_t)t * N, dir ? dir + (size_t)t * SPL_ACTIONS : nullptr, sc.dwords);
    w.sync();
    PROF_STAMP(A, t, 2, clock64());
    const int r = mcts_descend_tree<N, false>(w, A, t, 0, P, DESC_MAX_TERMINAL, max_levels, leaf_states + (size_t)t * MctsLay<N>::S, leaf_valids + (size_t)t * SPL_ACTIONS);
    PROF_STAMP(A, t, 3, clock64()); PROF_STAMP(A, t, 4, (long long)r * 1000 + A.trees[t].slot[0].path_len);
    if (RULES) {
        if (r == 2) rules_for_own_tree<N>(A, t, P.rules, sc.st, w.lane);
        PROF_STAMP(A, t, 9, clock64());
    }
    PROF_STAMP(A, t, 5, prof_globaltimer());
}

template <int N>
__global__ void __launch_bounds__(MW * 32) mcts_policy_kernel(MctsArena A, double temp, double* probs, double* q) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    WarpScratch sc = warp_scratch(warp);
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_policy_tree<N>(w, A, t, temp, probs + (size_t)t * SPL_ACTIONS, q + (size_t)t * N, sc.dwords);
}

template <int N>
__global__ void __launch_bounds__(MW * 32) mcts_sample_kernel(MctsArena A, MctsSearchParams P, double temp, const uint32_t* episodes, int16_t* actions, uint8_t* finished,
                                                              long long* counters) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    bool fin = false;
    const int a = mcts_sample_tree<N>(w, A, t, P, temp, episodes ? episodes[t] : 0u, &fin);
    if (w.lane == 0) {
        actions[t] = (int16_t)(fin ? a : -1);
        finished[t] = fin ? 1 : 0;
        if (counters && fin) {
            atomicAdd(reinterpret_cast<unsigned long long*>(&counters[0]), (unsigned long long)A.trees[t].sims_done);
            atomicAdd(reinterpret_cast<unsigned long long*>(&counters[1]), 1ull);
        }
    }
}

__global__ void __launch_bounds__(MW * 32) mcts_stats_kernel(MctsArena A, int32_t* nsa, double* qsa, float* ps, int32_t* info) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_root_stats_tree(w, A, t, nsa ? nsa + (size_t)t * SPL_ACTIONS : nullptr, qsa ? qsa + (size_t)t * SPL_ACTIONS : nullptr,
                         ps ? ps + (size_t)t * SPL_ACTIONS : nullptr, info ? info + (size_t)t * 16 : nullptr);
}

__global__ void __launch_bounds__(MW * 32) mcts_clean_kernel(MctsArena A, int max_nodes, int gc_reachable) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_clean_tree(w, A, t, max_nodes, gc_reachable);
}

// every page of the pool into the free ring (page 0 is never handed out: record offset 0 means "none")
__global__ void mcts_pool_init_kernel(MctsArena A) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.n_pool_pages) A.fq_slots[i] = i + 1u < A.n_pool_pages ? i + 1u : 0u;
    if (i == 0u) {
        A.fq_ctl[0] = 0; A.fq_ctl[1] = (int32_t)A.n_pool_pages - 1; A.fq_ctl[2] = (int32_t)A.n_pool_pages - 1; A.fq_ctl[3] = (int32_t)A.n_pool_pages - 1;
    }
}

__global__ void __launch_bounds__(MW * 32) mcts_reset_kernel(MctsArena A, const uint8_t* tree_select) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    if (tree_select && !tree_select[t]) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    if (!tree_select) {   // a full reset re-fills the free ring afterwards (mcts_pool_init_kernel): nothing to push page by page
        if (w.lane == 0) A.trees[t].n_pages = 0;
        __syncwarp();
    }
    mcts_clear_tree(w, A, t);
    if (w.lane == 0 && !tree_select) { A.trees[t].nn_calls = 0; A.trees[t].resets = 0; A.trees[t].compactions = 0; A.trees[t].truncated = 0; }   // a full reset also clears the statistics
}

template <int N>
__global__ void __launch_bounds__(MW * 32) mcts_fixed_net_kernel(const int8_t* states, const uint8_t* valids, int n_rows, float* pi, float* v) {
    const int warp = threadIdx.x >> 5, r = blockIdx.x * MW + warp;
    WarpScratch sc = warp_scratch(warp);
    if (r >= n_rows) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_fixed_net_row<N>(w, states + (size_t)r * MctsLay<N>::S, valids + (size_t)r * SPL_ACTIONS, pi + (size_t)r * SPL_ACTIONS, v + (size_t)r * N,
                          sc.words);
}

// ==========================================================================================
// host side of the C ABI
// ==========================================================================================
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct ArenaPlan {
    int sp, cp, hcap, max_depth, max_pages;
    uint32_t n_pool_pages;
    size_t off_pool, off_fq, off_ctl, off_tpages, off_htab, off_trees, off_path, off_sstate, off_smask, off_ses, off_sended, off_lsrc, total;
};
static ArenaPlan plan_arena(int n, int T, int node_limit, size_t pool_bytes, int K) {
    ArenaPlan p;
    p.sp = (7 * (32 + 10 * n + n * n) + 15) / 16 * 16;
    p.cp = n == 2 ? MctsCLay<2>::CP : n == 3 ? MctsCLay<3>::CP : MctsCLay<4>::CP;
    p.hcap = 64;
    while (p.hcap < 2 * node_limit) p.hcap *= 2;
    p.max_depth = 62 * n + 8;
    const size_t page_bytes = (size_t)MCTS_PAGE_UNITS * MCTS_UNIT;
    // at least one page per tree in flight plus the reserved page 0
    size_t pages = (pool_bytes + page_bytes - 1) / page_bytes;
    if (pages < (size_t)T + 8) pages = (size_t)T + 8;
    p.n_pool_pages = (uint32_t)pages;
    // per-tree page list: room for node_limit records of 40 edges, twice (the copy of a compaction)
    size_t half = ((size_t)node_limit * (size_t)(32 + p.cp + 24 * 40) + page_bytes - 1) / page_bytes + 2;
    if (half > pages) half = pages;
    p.max_pages = (int)(2 * half);
    size_t o = 0;
    p.off_pool = o;   o = align_up(o + pages * page_bytes, 256);
    p.off_fq = o;     o = align_up(o + pages * 4, 256);
    p.off_ctl = o;    o = align_up(o + 64, 256);
    p.off_tpages = o; o = align_up(o + (size_t)T * p.max_pages * 4, 256);
    p.off_htab = o;   o = align_up(o + (size_t)T * p.hcap * 4, 256);
    p.off_trees = o;  o = align_up(o + (size_t)T * sizeof(MctsTree), 256);
    const size_t rows = (size_t)T * K;      // (tree, in-flight slot)
    p.off_path = o;   o = align_up(o + rows * p.max_depth * 8, 256);
    p.off_sstate = o; o = align_up(o + rows * p.sp, 256);
    p.off_smask = o;  o = align_up(o + rows * 13 * 4, 256);
    p.off_ses = o;    o = align_up(o + rows * 16, 256);
    p.off_sended = o; o = align_up(o + rows, 256);
    p.off_lsrc = o;   o = align_up(o + rows, 256);
    p.total = o;
    return p;
}

#define ENTER_M(m)                                                        \
    if (!(m)) return spl_fail_(SPL_E_ARG, "null mcts handle");            \
    CU(cudaSetDevice((m)->ctx->device));                                  \
    cudaStream_t st = (cudaStream_t)stream;                               \
    const int grid = ((m)->A.n_trees + MW - 1) / MW;

extern "C" {

size_t spl_mcts_record_bytes(int n_players, int n_edges) {
    if (n_players < 2 || n_players > 4 || n_edges < 0) return 0;
    const int cp = n_players == 2 ? MctsCLay<2>::CP : n_players == 3 ? MctsCLay<3>::CP : MctsCLay<4>::CP;
    return (size_t)(32 + cp + 24 * n_edges + 31) / 32 * 32;
}

size_t spl_mcts_arena_bytes(int n_players, int n_trees, int node_limit, size_t pool_bytes, int leaves_per_tree) {
    if (n_players < 2 || n_players > 4 || n_trees <= 0 || node_limit <= 0 || leaves_per_tree < 1 || leaves_per_tree > MCTS_KMAX) return 0;
    return plan_arena(n_players, n_trees, node_limit, pool_bytes, leaves_per_tree).total;
}

int spl_mcts_create(spl_ctx* ctx, int n_trees, int node_limit, size_t pool_bytes, int leaves_per_tree, void* arena, size_t arena_bytes, spl_mcts** out) {
    if (!ctx || !out || !arena || n_trees <= 0 || node_limit < 4) return spl_fail_(SPL_E_ARG, "spl_mcts_create: bad argument");
    if (leaves_per_tree < 1 || leaves_per_tree > MCTS_KMAX) return spl_fail_(SPL_E_ARG, "spl_mcts_create: leaves_per_tree must be 1..4");
    if (node_limit > (1 << 24)) return spl_fail_(SPL_E_ARG, "spl_mcts_create: node_limit too large");
    if (((uintptr_t)arena & 255u) != 0) return spl_fail_(SPL_E_ARG, "spl_mcts_create: arena must be 256-byte aligned");
    const ArenaPlan p = plan_arena(ctx->n, n_trees, node_limit, pool_bytes, leaves_per_tree);
    if (arena_bytes < p.total) return spl_fail_(SPL_E_ARG, "spl_mcts_create: arena smaller than spl_mcts_arena_bytes");
    if ((size_t)p.n_pool_pages * MCTS_PAGE_UNITS > 0xFFFFFFFFull) return spl_fail_(SPL_E_ARG, "spl_mcts_create: pool larger than 128 GB");
    static_assert(sizeof(MctsNode) == 32 && sizeof(MctsPN) == 8 && sizeof(MctsCA) == 8 && sizeof(MctsTree) == 192, "arena record sizes");
    spl_mcts* m = new spl_mcts;
    m->ctx = ctx;
    char* base = (char*)arena;
    m->A.n_trees = n_trees; m->A.node_limit = node_limit; m->A.hcap = p.hcap; m->A.sp = p.sp; m->A.cp = p.cp; m->A.max_depth = p.max_depth;
    m->A.max_pages = p.max_pages; m->A.n_pool_pages = p.n_pool_pages; m->A.n_slots = leaves_per_tree;
    m->A.pool = (uint8_t*)(base + p.off_pool);
    m->A.fq_slots = (uint32_t*)(base + p.off_fq);
    m->A.fq_ctl = (int32_t*)(base + p.off_ctl);
    m->A.tree_pages = (uint32_t*)(base + p.off_tpages);
    m->A.htab = (uint32_t*)(base + p.off_htab);
    m->A.trees = (MctsTree*)(base + p.off_trees);
    m->A.path = (uint32_t*)(base + p.off_path);
    m->A.stage_state = (int8_t*)(base + p.off_sstate);
    m->A.stage_mask = (uint32_t*)(base + p.off_smask);
    m->A.stage_es = (float*)(base + p.off_ses);
    m->A.stage_ended = (uint8_t*)(base + p.off_sended);
    m->A.leaf_src = (uint8_t*)(base + p.off_lsrc);
    m->A.prof = nullptr;
    m->P.cpuct = 1.0; m->P.fpu = 0.0; m->P.temperature0 = 1.0; m->P.dirichlet_alpha = 0.3; m->P.seed = 0; m->P.game_base = 0;
    m->P.rules = ctx->rules;
    m->gc_reachable = 0; m->rounds = 1; m->max_levels = 1 << 20;
    m->episodes = nullptr;
    m->rules_tpw = 8;
    if (const char* e = getenv("SPL_MCTS_RULES_TPW")) {   // tuning hook (4, 8 or 16)
        const int v = atoi(e);
        if (v == 4 || v == 8 || v == 16) m->rules_tpw = v;
    }
    // rules step inside the descent kernel: one launch boundary less on the critical path of a wave, but the step then runs on one
    // lane per warp - right while every tree's warp is resident at once (latency-bound, measured 102 -> 98 us per wave at 4096
    // trees), wrong in the throughput regime (65,536 trees: 32x the issue slots of mcts_rules_kernel)
    m->fuse_rules = n_trees <= 6144 && leaves_per_tree == 1;
    if (const char* e = getenv("SPL_MCTS_FUSE_RULES")) m->fuse_rules = atoi(e) != 0 && leaves_per_tree == 1;   // tuning hook
    m->pdl = 1;
    if (const char* e = getenv("SPL_MCTS_PDL")) m->pdl = atoi(e) != 0;   // tuning hook
    m->side = nullptr; m->ev_fork = nullptr; m->ev_join = nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        delete m;
        return spl_fail_(SPL_E_CUDA, "spl_mcts_create: stream / event creation failed");
    }
    *out = m;
    return SPL_OK;   // the caller resets the arena (spl_mcts_reset with tree_select == NULL) before the first begin
}

void spl_mcts_destroy(spl_mcts* m) {
    if (!m) return;
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    if (m->side) cudaStreamDestroy(m->side);
    delete m;
}

int spl_mcts_set_params(spl_mcts* m, const spl_mcts_params* p) {
    if (!m || !p) return spl_fail_(SPL_E_ARG, "spl_mcts_set_params: null argument");
    if (!(p->temperature0 > 0.0) || !(p->dirichlet_alpha > 0.0)) return spl_fail_(SPL_E_ARG, "spl_mcts_set_params: bad value");
    m->P.cpuct = p->cpuct; m->P.fpu = p->fpu; m->P.temperature0 = p->temperature0; m->P.dirichlet_alpha = p->dirichlet_alpha;
    m->P.seed = p->seed; m->P.game_base = p->game_base;
    m->gc_reachable = p->gc_reachable ? 1 : 0;
    m->rounds = p->rounds < 1 ? 1 : (p->rounds > 8 ? 8 : p->rounds);
    m->max_levels = p->max_levels < 1 ? (1 << 20) : p->max_levels;
    return SPL_OK;
}

int spl_mcts_reset(spl_mcts* m, const uint8_t* tree_select, void* stream) {
    ENTER_M(m);
    mcts_reset_kernel<<<grid, MW * 32, 0, st>>>(m->A, tree_select);
    if (!tree_select) mcts_pool_init_kernel<<<(m->A.n_pool_pages + 255) / 256, 256, 0, st>>>(m->A);
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_clean(spl_mcts* m, int fill_percent, void* stream) {
    ENTER_M(m);
    if (fill_percent < 0 || fill_percent > 100) return spl_fail_(SPL_E_ARG, "spl_mcts_clean: bad argument");
    mcts_clean_kernel<<<grid, MW * 32, 0, st>>>(m->A, (int)((long long)m->A.node_limit * fill_percent / 100), m->gc_reachable);
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_begin(spl_mcts* m, const int8_t* roots, const int32_t* sims, const uint8_t* move_flags, const uint8_t* tree_select,
                   const double* dir_values, void* stream) {
    ENTER_M(m);
    if (!roots || !sims) return spl_fail_(SPL_E_ARG, "spl_mcts_begin: bad argument");
    m->P.rules = m->ctx->rules;
    DISPATCH_N(m->ctx->n, mcts_begin_kernel<N><<<grid, MW * 32, 0, st>>>(m->A, m->P, roots, sims, move_flags, tree_select, dir_values, m->episodes, m->gc_reachable));
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_select(spl_mcts* m, int8_t* leaf_states, uint8_t* leaf_valids, uint8_t* leaf_flags, int32_t* counters, void* stream) {
    ENTER_M(m);
    if (!leaf_states || !leaf_valids || !leaf_flags) return spl_fail_(SPL_E_ARG, "spl_mcts_select: bad argument");
    const SplRules rules = m->ctx->rules;
    const bool vl = m->A.n_slots > 1;
    DISPATCH_N(m->ctx->n, {
        for (int r = 0; r < m->rounds; r++) {
            int32_t* cnt = r == m->rounds - 1 ? counters : nullptr;
            if (vl) mcts_descend_kernel<N, true><<<grid, MW * 32, 0, st>>>(m->A, m->P, m->max_levels, leaf_states, leaf_valids);
            else mcts_descend_kernel<N, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, m->max_levels, leaf_states, leaf_valids);
            CU(launch_rules<N>(m->A, rules, m->rules_tpw, st));
            if (vl) mcts_attach_kernel<N, true><<<grid, MW * 32, 0, st>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, cnt, true);
            else mcts_attach_kernel<N, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, cnt, true);
        }
    });
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_expand_select(spl_mcts* m, const float* pi, const float* v, const double* dir_values, int8_t* leaf_states, uint8_t* leaf_valids,
                           uint8_t* leaf_flags, int32_t* counters, void* stream) {
    ENTER_M(m);
    if (!pi || !v || !leaf_states || !leaf_valids || !leaf_flags) return spl_fail_(SPL_E_ARG, "spl_mcts_expand_select: bad argument");
    const SplRules rules = m->ctx->rules;
    const bool vl = m->A.n_slots > 1;
    DISPATCH_N(m->ctx->n, {
        for (int r = 0; r < m->rounds; r++) {
            int32_t* cnt = r == m->rounds - 1 ? counters : nullptr;
            if (r == 0 && vl) mcts_expand_descend_kernel<N, false, true><<<grid, MW * 32, 0, st>>>(m->A, m->P, pi, v, dir_values, m->max_levels, leaf_states, leaf_valids);
            else if (r == 0) mcts_expand_descend_kernel<N, false, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, pi, v, dir_values, m->max_levels, leaf_states, leaf_valids);
            else if (vl) mcts_descend_kernel<N, true><<<grid, MW * 32, 0, st>>>(m->A, m->P, m->max_levels, leaf_states, leaf_valids);
            else mcts_descend_kernel<N, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, m->max_levels, leaf_states, leaf_valids);
            CU(launch_rules<N>(m->A, rules, m->rules_tpw, st));
            if (vl) mcts_attach_kernel<N, true><<<grid, MW * 32, 0, st>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, cnt, true);
            else mcts_attach_kernel<N, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, cnt, true);
        }
    });
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_wave_nnet(spl_mcts* m, const void* nnet_blob, float* pi, float* v, const double* dir_values, int8_t* leaf_states,
                       uint8_t* leaf_valids, uint8_t* leaf_flags, int32_t* counters, void* stream) {
    ENTER_M(m);
    if (!nnet_blob || !pi || !v || !leaf_states || !leaf_valids || !leaf_flags) return spl_fail_(SPL_E_ARG, "spl_mcts_wave_nnet: bad argument");
    const SplRules rules = m->ctx->rules;
    const bool vl = m->A.n_slots > 1;
    const int n_rows = m->A.n_trees * m->A.n_slots;
    DISPATCH_N(m->ctx->n, {
        m->P.rules = rules;
        if (m->fuse_rules && m->pdl && m->rounds <= 1) {
            // stream: [expand + descend + rules] -> [network] -> next wave, both launched as programmatic dependents of their
            // predecessor (their blocks become resident - and the network does its set-up: parameters, TMEM, first weight tiles -
            // while the predecessor drains; griddepcontrol.wait guards the first dependent read); the attach kernel runs on the
            // side stream between this wave's descent and the next one's expansion
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(MW * 32); cfg.dynamicSmemBytes = 0; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
            CU(cudaLaunchKernelEx(&cfg, mcts_expand_descend_kernel<N, true, false>, m->A, m->P, (const float*)pi, (const float*)v, dir_values, m->max_levels,
                                  leaf_states, leaf_valids));
            CU(cudaEventRecord(m->ev_fork, st));
            CU(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
            mcts_attach_kernel<N, false><<<grid, MW * 32, 0, m->side>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, counters, false);
            CU(cudaEventRecord(m->ev_join, m->side));
            const int rc = spl_nnet_forward_rows_(m->ctx, nnet_blob, leaf_states, leaf_valids, m->A.leaf_src, m->A.stage_state, m->A.sp, m->A.stage_mask,
                                                  m->A.n_trees, m->A.n_trees, pi, v, st, true);
            if (rc != SPL_OK) return rc;
            CU(cudaStreamWaitEvent(st, m->ev_join, 0));
            CU(cudaGetLastError());
            return SPL_OK;
        }
        if (m->fuse_rules && m->rounds <= 1) {
            mcts_expand_descend_kernel<N, true, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, pi, v, dir_values, m->max_levels, leaf_states, leaf_valids);
        } else {
            if (vl) mcts_expand_descend_kernel<N, false, true><<<grid, MW * 32, 0, st>>>(m->A, m->P, pi, v, dir_values, m->max_levels, leaf_states, leaf_valids);
            else mcts_expand_descend_kernel<N, false, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, pi, v, dir_values, m->max_levels, leaf_states, leaf_valids);
            CU(launch_rules<N>(m->A, rules, m->rules_tpw, st));
            // rounds > 1 (lock-step callers): a tree whose descent ran into a transposition or backed a terminal value up gets another
            // (attach, descend, rules) pass before the network runs, so that the wave still ends with a leaf for it. Scheduling only:
            // every tree runs the same sequential search
            for (int r = 1; r < m->rounds; r++) {
                if (vl) mcts_attach_kernel<N, true><<<grid, MW * 32, 0, st>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, nullptr, false);
                else mcts_attach_kernel<N, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, nullptr, false);
                if (vl) mcts_descend_kernel<N, true><<<grid, MW * 32, 0, st>>>(m->A, m->P, m->max_levels, leaf_states, leaf_valids);
                else mcts_descend_kernel<N, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, m->max_levels, leaf_states, leaf_valids);
                CU(launch_rules<N>(m->A, rules, m->rules_tpw, st));
            }
        }
        // fork: the network reads the rows the descent (leaf rows) or the rules kernel (staging rows) just wrote, while the
        // attach kernel links the new children into the trees; join before the next wave's expansion reads pi / v
        CU(cudaEventRecord(m->ev_fork, st));
        CU(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
        const int rc = spl_nnet_forward_rows_(m->ctx, nnet_blob, leaf_states, leaf_valids, m->A.leaf_src, m->A.stage_state, m->A.sp, m->A.stage_mask,
                                              n_rows, n_rows, pi, v, m->side, false);
        if (rc != SPL_OK) return rc;
        CU(cudaEventRecord(m->ev_join, m->side));
        if (vl) mcts_attach_kernel<N, true><<<grid, MW * 32, 0, st>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, counters, false);
        else mcts_attach_kernel<N, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, counters, false);
        CU(cudaStreamWaitEvent(st, m->ev_join, 0));
    });
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_expand(spl_mcts* m, const float* pi, const float* v, const double* dir_values, void* stream) {
    ENTER_M(m);
    if (!pi || !v) return spl_fail_(SPL_E_ARG, "spl_mcts_expand: bad argument");
    if (m->A.n_slots > 1) { DISPATCH_N(m->ctx->n, (mcts_expand_kernel<N, true><<<grid, MW * 32, 0, st>>>(m->A, m->P, pi, v, dir_values))); }
    else { DISPATCH_N(m->ctx->n, (mcts_expand_kernel<N, false><<<grid, MW * 32, 0, st>>>(m->A, m->P, pi, v, dir_values))); }
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_policy(spl_mcts* m, double temp, double* probs, double* q, void* stream) {
    ENTER_M(m);
    if (!probs || !q || temp < 0.0) return spl_fail_(SPL_E_ARG, "spl_mcts_policy: bad argument");
    DISPATCH_N(m->ctx->n, mcts_policy_kernel<N><<<grid, MW * 32, 0, st>>>(m->A, temp, probs, q));
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_sample_moves(spl_mcts* m, double temp, const uint32_t* episodes, int16_t* actions, uint8_t* finished, long long* counters, void* stream) {
    ENTER_M(m);
    if (!actions || !finished || temp < 0.0) return spl_fail_(SPL_E_ARG, "spl_mcts_sample_moves: bad argument");
    DISPATCH_N(m->ctx->n, mcts_sample_kernel<N><<<grid, MW * 32, 0, st>>>(m->A, m->P, temp, episodes, actions, finished, counters));
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_root_stats(spl_mcts* m, int32_t* nsa, double* qsa, float* ps, int32_t* info, void* stream) {
    ENTER_M(m);
    mcts_stats_kernel<<<grid, MW * 32, 0, st>>>(m->A, nsa, qsa, ps, info);
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_set_episodes(spl_mcts* m, const uint32_t* episodes) {
    if (!m) return spl_fail_(SPL_E_ARG, "null mcts handle");
    m->episodes = episodes;
    return SPL_OK;
}

int spl_mcts_pool_stats(spl_mcts* m, int32_t* out4, void* stream) {   /* host buffer: pages in the pool, free now, fewest free so far, bytes per page */
    ENTER_M(m);
    (void)grid;
    if (!out4) return spl_fail_(SPL_E_ARG, "spl_mcts_pool_stats: bad argument");
    int32_t ctl[4];
    CU(cudaMemcpyAsync(ctl, m->A.fq_ctl, sizeof ctl, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    out4[0] = (int32_t)m->A.n_pool_pages - 1; out4[1] = ctl[2]; out4[2] = ctl[3]; out4[3] = (int32_t)(MCTS_PAGE_UNITS * MCTS_UNIT);
    return SPL_OK;
}

int spl_mcts_debug_profile(spl_mcts* m, long long* stamps) {   /* diagnostics only: int64[T][16] device buffer, NULL = off */
    if (!m) return spl_fail_(SPL_E_ARG, "null mcts handle");
    m->A.prof = stamps;
    return SPL_OK;
}

int spl_mcts_fixed_net(spl_ctx* c, const int8_t* states, const uint8_t* valids, int n_rows, float* pi, float* v, void* stream) {
    if (!c) return spl_fail_(SPL_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (!states || !valids || !pi || !v || n_rows <= 0) return spl_fail_(SPL_E_ARG, "spl_mcts_fixed_net: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_N(c->n, mcts_fixed_net_kernel<N><<<(n_rows + MW - 1) / MW, MW * 32, 0, st>>>(states, valids, n_rows, pi, v));
    CU(cudaGetLastError());
    return SPL_OK;
}

}   // extern "C"
