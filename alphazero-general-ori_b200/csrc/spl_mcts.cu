// libsplendor_b200.so - MCTS tree arena kernels for sm_100a and their C ABI (include/splendor_b200.h).
// The per-tree logic lives in spl_mcts.cuh; this file is launch plumbing. A selection wave is three kernels:
//   mcts_descend_kernel  one warp per tree, light (PUCT pick + path only, no rules code -> every tree resident at once)
//   mcts_rules_kernel    one LANE per tree, 32 trees per warp on a shared-memory tile like the environment kernels:
//                        make_move + swap_players + getGameEnded + getValidMoves of the child of every pending edge
//   mcts_attach_kernel   one warp per tree: hash, dictionary lookup / insertion, edge allocation, leaf hand-over
// A tree whose new edge led into a node it already holds (transposition) or into a terminal node carries on in the next
// wave; `rounds` > 1 repeats the three kernels inside one wave instead (measured: not worth the extra straggler-bound launches).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "spl_internal.h"
#include "spl_mcts.cuh"

#define MW 4                 // warps (trees) per CTA
#define MSP 640              // per-warp state scratch (>= MctsLay<4>::SP = 624)

struct spl_mcts {
    spl_ctx* ctx;
    MctsArena A;
    MctsSearchParams P;
    int edge_reserve, gc_reachable, rounds, max_levels;
};

struct WarpScratch {
    int8_t* st;
    uint32_t* words;
    double* dwords;
};

__device__ __forceinline__ WarpScratch warp_scratch(int warp) {
    __shared__ __align__(16) int8_t s_state[MW][MSP];
    __shared__ uint32_t s_words[MW][24];
    __shared__ double s_dwords[MW][4];
    WarpScratch s;
    s.st = s_state[warp]; s.words = s_words[warp]; s.dwords = s_dwords[warp];
    return s;
}

template <int N>
__global__ void __launch_bounds__(MW * 32) mcts_begin_kernel(MctsArena A, MctsSearchParams P, const int8_t* roots, const int32_t* sims,
                                                             const uint8_t* move_flags, const uint8_t* tree_select, const double* dir,
                                                             int edge_reserve, int gc_reachable) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    WarpScratch sc = warp_scratch(warp);
    if (t >= A.n_trees) return;
    if (tree_select && !tree_select[t]) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_begin_tree<N>(w, A, t, P, roots + (size_t)t * MctsLay<N>::S, sims[t], move_flags ? (uint32_t)move_flags[t] : 0u, edge_reserve,
                       gc_reachable, dir ? dir + (size_t)t * SPL_ACTIONS : nullptr, sc.st, sc.words, sc.dwords);
}

template <int N>
__global__ void __launch_bounds__(MW * 32, 8) mcts_descend_kernel(MctsArena A, MctsSearchParams P, int max_levels, int8_t* leaf_states,
                                                                  uint8_t* leaf_valids) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_descend_tree<N>(w, A, t, P, 1, max_levels, leaf_states + (size_t)t * MctsLay<N>::S, leaf_valids + (size_t)t * SPL_ACTIONS);
}

// shared-memory tile accessor of the rules kernel: cell (row, col) of this lane's tree = byte [(7 row + col) * 32 + lane]
struct TreeTileAcc {
    int8_t* b;
    __device__ __forceinline__ int get(int row, int col) const { return b[(row * 7 + col) * 32]; }
    __device__ __forceinline__ void set(int row, int col, int v) { b[(row * 7 + col) * 32] = (int8_t)v; }
};
#define RW 2   // warps (tiles of 32 trees) per CTA of the rules kernel

template <int N>
__global__ void __launch_bounds__(RW * 32) mcts_rules_kernel(MctsArena A, SplRules rules) {
    typedef MctsLay<N> ML;
    extern __shared__ __align__(16) int8_t tile_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = (blockIdx.x * RW + warp) * 32 + lane;
    int8_t* col = tile_smem + (size_t)warp * ML::S * 32 + lane;
    bool pending = false;
    int parent = 0, action = 0;
    if (t < A.n_trees) {
        const MctsTree* T = A.trees + t;
        const int pe = T->pend_edge;
        if (pe >= 0 && T->leaf < 0) {
            pending = true;
            parent = T->pend_parent;
            action = (int)A.edges[(size_t)t * A.ecap + pe].action;
        }
    }
    if (!__any_sync(0xffffffffu, pending)) return;
    if (pending) {
        const uint4* src = reinterpret_cast<const uint4*>(A.states + ((size_t)t * A.cap + parent) * A.sp);
#pragma unroll 5
        for (int i = 0; i < ML::SP / 16; i++) {   // own state -> own column (every store of the warp hits one cell row: conflict-free)
            const uint4 v = __ldg(src + i);
            const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int cell = 16 * i + j;
                if (cell < ML::S) col[cell * 32] = (int8_t)((wv[j >> 2] >> (8 * (j & 3))) & 0xFFu);
            }
        }
        TreeTileAcc s{col};
        float es[N];
        uint32_t m[SPL_MASK_WORDS];
        const bool ended = mcts_rules_core<N>(s, action, rules, es, m);
        uint4* dst = reinterpret_cast<uint4*>(A.stage_state + (size_t)t * A.sp);
#pragma unroll 5
        for (int i = 0; i < ML::SP / 16; i++) {
            uint32_t wv[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int cell = 16 * i + j;
                if (cell < ML::S) wv[j >> 2] |= (uint32_t)(uint8_t)col[cell * 32] << (8 * (j & 3));
            }
            dst[i] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
#pragma unroll
        for (int i = 0; i < SPL_MASK_WORDS; i++) A.stage_mask[(size_t)i * A.n_trees + t] = m[i];
#pragma unroll
        for (int i = 0; i < 4; i++) A.stage_es[(size_t)t * 4 + i] = i < N ? es[i] : 0.f;
        A.stage_ended[t] = ended ? 1 : 0;
    }
}

template <int N>
__global__ void __launch_bounds__(MW * 32, 8) mcts_attach_kernel(MctsArena A, MctsSearchParams P, int8_t* leaf_states, uint8_t* leaf_valids,
                                                                 uint8_t* leaf_flags, int32_t* counters) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    const int leaf = mcts_attach_tree<N>(w, A, t, P, A.stage_state + (size_t)t * A.sp, A.stage_ended[t] != 0, A.stage_es + (size_t)t * 4,
                                         A.stage_mask + t, A.n_trees, leaf_states + (size_t)t * MctsLay<N>::S, leaf_valids + (size_t)t * SPL_ACTIONS);
    if (w.lane == 0) {
        leaf_flags[t] = (uint8_t)leaf;
        if (counters) {   // last round of the wave
            const MctsTree* T = A.trees + t;
            if (leaf) atomicAdd(&counters[0], 1);
            if (T->status == 0u && (leaf || T->pend_edge >= 0 || T->sims_done < T->sims_target)) atomicAdd(&counters[1], 1);
        }
    }
}

template <int N>
__global__ void __launch_bounds__(MW * 32, 8) mcts_expand_kernel(MctsArena A, MctsSearchParams P, const float* pi, const float* v, const double* dir) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    WarpScratch sc = warp_scratch(warp);
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_expand_tree<N>(w, A, t, P, pi + (size_t)t * SPL_ACTIONS, v + (size_t)t * N, dir ? dir + (size_t)t * SPL_ACTIONS : nullptr, sc.dwords);
}

// expansion of the previous wave's leaf and the next descent of the same tree in one launch (both are warp-per-tree)
template <int N>
__global__ void __launch_bounds__(MW * 32, 8) mcts_expand_descend_kernel(MctsArena A, MctsSearchParams P, const float* pi, const float* v, const double* dir,
                                                                         int max_levels, int8_t* leaf_states, uint8_t* leaf_valids) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    WarpScratch sc = warp_scratch(warp);
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_expand_tree<N>(w, A, t, P, pi + (size_t)t * SPL_ACTIONS, v + (size_t)t * N, dir ? dir + (size_t)t * SPL_ACTIONS : nullptr, sc.dwords);
    w.sync();
    mcts_descend_tree<N>(w, A, t, P, 1, max_levels, leaf_states + (size_t)t * MctsLay<N>::S, leaf_valids + (size_t)t * SPL_ACTIONS);
}

template <int N>
__global__ void __launch_bounds__(MW * 32) mcts_policy_kernel(MctsArena A, double temp, double* probs, double* q) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    WarpScratch sc = warp_scratch(warp);
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_policy_tree<N>(w, A, t, temp, probs + (size_t)t * SPL_ACTIONS, q + (size_t)t * N, sc.dwords);
}

__global__ void __launch_bounds__(MW * 32) mcts_stats_kernel(MctsArena A, int32_t* nsa, double* qsa, float* ps, int32_t* info) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_root_stats_tree(w, A, t, nsa ? nsa + (size_t)t * SPL_ACTIONS : nullptr, qsa ? qsa + (size_t)t * SPL_ACTIONS : nullptr,
                         ps ? ps + (size_t)t * SPL_ACTIONS : nullptr, info ? info + (size_t)t * 16 : nullptr);
}

__global__ void __launch_bounds__(MW * 32) mcts_clean_kernel(MctsArena A, int max_nodes, int max_edges, int gc_reachable) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
    mcts_clean_tree(w, A, t, max_nodes, max_edges, gc_reachable);
}

__global__ void __launch_bounds__(MW * 32) mcts_reset_kernel(MctsArena A, const uint8_t* tree_select) {
    const int warp = threadIdx.x >> 5, t = blockIdx.x * MW + warp;
    if (t >= A.n_trees) return;
    if (tree_select && !tree_select[t]) return;
    MctsWarp w{(int)(threadIdx.x & 31)};
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
    int sp, hcap, max_depth;
    size_t off_states, off_nodes, off_edges, off_htab, off_trees, off_path, off_sstate, off_smask, off_ses, off_sended, total;
};
static ArenaPlan plan_arena(int n, int T, int cap, int ecap) {
    ArenaPlan p;
    p.sp = (7 * (32 + 10 * n + n * n) + 15) / 16 * 16;
    p.hcap = 64;
    while (p.hcap < 2 * cap) p.hcap *= 2;
    p.max_depth = 62 * n + 8;
    size_t o = 0;
    p.off_states = o; o = align_up(o + (size_t)T * cap * p.sp, 256);
    p.off_nodes = o;  o = align_up(o + (size_t)T * cap * sizeof(MctsNode), 256);
    p.off_edges = o;  o = align_up(o + (size_t)T * ecap * sizeof(MctsEdge), 256);
    p.off_htab = o;   o = align_up(o + (size_t)T * p.hcap * 4, 256);
    p.off_trees = o;  o = align_up(o + (size_t)T * sizeof(MctsTree), 256);
    p.off_path = o;   o = align_up(o + (size_t)T * p.max_depth * 8, 256);
    p.off_sstate = o; o = align_up(o + (size_t)T * p.sp, 256);
    p.off_smask = o;  o = align_up(o + (size_t)T * 13 * 4, 256);
    p.off_ses = o;    o = align_up(o + (size_t)T * 16, 256);
    p.off_sended = o; o = align_up(o + (size_t)T, 256);
    p.total = o;
    return p;
}

#define ENTER_M(m)                                                        \
    if (!(m)) return spl_fail_(SPL_E_ARG, "null mcts handle");            \
    CU(cudaSetDevice((m)->ctx->device));                                  \
    cudaStream_t st = (cudaStream_t)stream;                               \
    const int grid = ((m)->A.n_trees + MW - 1) / MW;

extern "C" {

size_t spl_mcts_arena_bytes(int n_players, int n_trees, int node_cap, int edge_cap) {
    if (n_players < 2 || n_players > 4 || n_trees <= 0 || node_cap <= 0 || edge_cap <= 0) return 0;
    return plan_arena(n_players, n_trees, node_cap, edge_cap).total;
}

int spl_mcts_create(spl_ctx* ctx, int n_trees, int node_cap, int edge_cap, void* arena, size_t arena_bytes, spl_mcts** out) {
    if (!ctx || !out || !arena || n_trees <= 0 || node_cap < 4 || edge_cap < 4) return spl_fail_(SPL_E_ARG, "spl_mcts_create: bad argument");
    if (node_cap > (1 << 24)) return spl_fail_(SPL_E_ARG, "spl_mcts_create: node_cap too large");
    if (((uintptr_t)arena & 255u) != 0) return spl_fail_(SPL_E_ARG, "spl_mcts_create: arena must be 256-byte aligned");
    const ArenaPlan p = plan_arena(ctx->n, n_trees, node_cap, edge_cap);
    if (arena_bytes < p.total) return spl_fail_(SPL_E_ARG, "spl_mcts_create: arena smaller than spl_mcts_arena_bytes");
    static_assert(sizeof(MctsNode) == 32 && sizeof(MctsEdge) == 32 && sizeof(MctsTree) == 96, "arena record sizes");
    spl_mcts* m = new spl_mcts;
    m->ctx = ctx;
    char* base = (char*)arena;
    m->A.n_trees = n_trees; m->A.cap = node_cap; m->A.ecap = edge_cap; m->A.hcap = p.hcap; m->A.sp = p.sp; m->A.max_depth = p.max_depth;
    m->A.states = (int8_t*)(base + p.off_states);
    m->A.nodes = (MctsNode*)(base + p.off_nodes);
    m->A.edges = (MctsEdge*)(base + p.off_edges);
    m->A.htab = (uint32_t*)(base + p.off_htab);
    m->A.trees = (MctsTree*)(base + p.off_trees);
    m->A.path = (uint32_t*)(base + p.off_path);
    m->A.stage_state = (int8_t*)(base + p.off_sstate);
    m->A.stage_mask = (uint32_t*)(base + p.off_smask);
    m->A.stage_es = (float*)(base + p.off_ses);
    m->A.stage_ended = (uint8_t*)(base + p.off_sended);
    m->P.cpuct = 1.0; m->P.fpu = 0.0; m->P.temperature0 = 1.0; m->P.dirichlet_alpha = 0.3; m->P.seed = 0; m->P.game_base = 0;
    m->P.rules = ctx->rules;
    m->edge_reserve = 32; m->gc_reachable = 0; m->rounds = 1; m->max_levels = 1 << 20;
    *out = m;
    return SPL_OK;
}

void spl_mcts_destroy(spl_mcts* m) { delete m; }

int spl_mcts_set_params(spl_mcts* m, const spl_mcts_params* p) {
    if (!m || !p) return spl_fail_(SPL_E_ARG, "spl_mcts_set_params: null argument");
    if (!(p->temperature0 > 0.0) || !(p->dirichlet_alpha > 0.0) || p->edge_reserve < 1) return spl_fail_(SPL_E_ARG, "spl_mcts_set_params: bad value");
    m->P.cpuct = p->cpuct; m->P.fpu = p->fpu; m->P.temperature0 = p->temperature0; m->P.dirichlet_alpha = p->dirichlet_alpha;
    m->P.seed = p->seed; m->P.game_base = p->game_base;
    m->edge_reserve = p->edge_reserve; m->gc_reachable = p->gc_reachable ? 1 : 0;
    m->rounds = p->rounds < 1 ? 1 : (p->rounds > 8 ? 8 : p->rounds);
    m->max_levels = p->max_levels < 1 ? (1 << 20) : p->max_levels;
    return SPL_OK;
}

int spl_mcts_reset(spl_mcts* m, const uint8_t* tree_select, void* stream) {
    ENTER_M(m);
    mcts_reset_kernel<<<grid, MW * 32, 0, st>>>(m->A, tree_select);
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_clean(spl_mcts* m, int fill_percent, void* stream) {
    ENTER_M(m);
    if (fill_percent < 0 || fill_percent > 100) return spl_fail_(SPL_E_ARG, "spl_mcts_clean: bad argument");
    mcts_clean_kernel<<<grid, MW * 32, 0, st>>>(m->A, (int)((long long)m->A.cap * fill_percent / 100), (int)((long long)m->A.ecap * fill_percent / 100),
                                                m->gc_reachable);
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_begin(spl_mcts* m, const int8_t* roots, const int32_t* sims, const uint8_t* move_flags, const uint8_t* tree_select,
                   const double* dir_values, void* stream) {
    ENTER_M(m);
    if (!roots || !sims) return spl_fail_(SPL_E_ARG, "spl_mcts_begin: bad argument");
    m->P.rules = m->ctx->rules;
    DISPATCH_N(m->ctx->n, mcts_begin_kernel<N><<<grid, MW * 32, 0, st>>>(m->A, m->P, roots, sims, move_flags, tree_select, dir_values, m->edge_reserve, m->gc_reachable));
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_select(spl_mcts* m, int8_t* leaf_states, uint8_t* leaf_valids, uint8_t* leaf_flags, int32_t* counters, void* stream) {
    ENTER_M(m);
    if (!leaf_states || !leaf_valids || !leaf_flags) return spl_fail_(SPL_E_ARG, "spl_mcts_select: bad argument");
    const int tiles = (m->A.n_trees + 31) / 32;
    const SplRules rules = m->ctx->rules;
    DISPATCH_N(m->ctx->n, {
        const int smem = RW * MctsLay<N>::S * 32;
        auto rk = mcts_rules_kernel<N>;
        CU(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int r = 0; r < m->rounds; r++) {
            int32_t* cnt = r == m->rounds - 1 ? counters : nullptr;
            mcts_descend_kernel<N><<<grid, MW * 32, 0, st>>>(m->A, m->P, m->max_levels, leaf_states, leaf_valids);
            rk<<<(tiles + RW - 1) / RW, RW * 32, smem, st>>>(m->A, rules);
            mcts_attach_kernel<N><<<grid, MW * 32, 0, st>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, cnt);
        }
    });
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_expand_select(spl_mcts* m, const float* pi, const float* v, const double* dir_values, int8_t* leaf_states, uint8_t* leaf_valids,
                           uint8_t* leaf_flags, int32_t* counters, void* stream) {
    ENTER_M(m);
    if (!pi || !v || !leaf_states || !leaf_valids || !leaf_flags) return spl_fail_(SPL_E_ARG, "spl_mcts_expand_select: bad argument");
    const int tiles = (m->A.n_trees + 31) / 32;
    const SplRules rules = m->ctx->rules;
    DISPATCH_N(m->ctx->n, {
        const int smem = RW * MctsLay<N>::S * 32;
        auto rk = mcts_rules_kernel<N>;
        CU(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int r = 0; r < m->rounds; r++) {
            int32_t* cnt = r == m->rounds - 1 ? counters : nullptr;
            if (r == 0) mcts_expand_descend_kernel<N><<<grid, MW * 32, 0, st>>>(m->A, m->P, pi, v, dir_values, m->max_levels, leaf_states, leaf_valids);
            else mcts_descend_kernel<N><<<grid, MW * 32, 0, st>>>(m->A, m->P, m->max_levels, leaf_states, leaf_valids);
            rk<<<(tiles + RW - 1) / RW, RW * 32, smem, st>>>(m->A, rules);
            mcts_attach_kernel<N><<<grid, MW * 32, 0, st>>>(m->A, m->P, leaf_states, leaf_valids, leaf_flags, cnt);
        }
    });
    CU(cudaGetLastError());
    return SPL_OK;
}

int spl_mcts_expand(spl_mcts* m, const float* pi, const float* v, const double* dir_values, void* stream) {
    ENTER_M(m);
    if (!pi || !v) return spl_fail_(SPL_E_ARG, "spl_mcts_expand: bad argument");
    DISPATCH_N(m->ctx->n, mcts_expand_kernel<N><<<grid, MW * 32, 0, st>>>(m->A, m->P, pi, v, dir_values));
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

int spl_mcts_root_stats(spl_mcts* m, int32_t* nsa, double* qsa, float* ps, int32_t* info, void* stream) {
    ENTER_M(m);
    mcts_stats_kernel<<<grid, MW * 32, 0, st>>>(m->A, nsa, qsa, ps, info);
    CU(cudaGetLastError());
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
