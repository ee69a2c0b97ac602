// TEST-ONLY: compiles the tree-arena logic (csrc/spl_mcts.cuh) with g++ and a one-lane "warp" so that the search
// semantics can be checked against the golden fixtures in the GPU-less build container. Never loaded by the product.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
struct uint4 { uint32_t x, y, z, w; };
#include "../../alphazero-general-ori_b200/csrc/spl_mcts.cuh"

#define DISPATCH(n, CALL) switch (n) { case 2: { constexpr int N = 2; CALL; } break; case 3: { constexpr int N = 3; CALL; } break; default: { constexpr int N = 4; CALL; } }

struct HsMcts {
    int n;
    MctsArena A;
    MctsSearchParams P;
    int edge_reserve, gc_reachable;
    uint32_t episode;
    int clean_every, clean_gc, clean_counter;   // test hook: clean in the middle of a search every N driver steps
    int8_t* leaf_state;
    uint8_t* leaf_valid;
    float *pi, *v;
};

template <int N>
static void hm_step_rules(HsMcts* m, int8_t* st, bool* ended, float* es, uint32_t* mask) {
    const MctsSlot& T = m->A.trees[0].slot[0];
    MctsWarp w{0};
    mcts_decode<N>(w, mcts_cstate(m->A, T.pend_parent), st);
    AosAcc s{st};
    const int action = (int)mcts_edges(m->A, T.pend_parent, T.pend_edge >> 16).ca[T.pend_edge & 0xFFFF].action;
    *ended = mcts_rules_core<N>(s, action, m->P.rules, es, mask);
}

extern "C" {
// cap: node limit of the tree; pool_nodes: how many average-size records the page pool holds (0: cap)
HsMcts* hm_create(int n, int cap, int ecap, int limit, uint32_t rule_flags, double cpuct, double fpu, double temperature0, int edge_reserve, int gc_reachable) {
    HsMcts* m = (HsMcts*)calloc(1, sizeof *m);
    m->n = n;
    MctsArena& A = m->A;
    A.n_trees = 1; A.node_limit = cap; A.n_slots = 1;
    A.hcap = 64; while (A.hcap < 2 * cap) A.hcap *= 2;
    A.sp = (7 * (32 + 10 * n + n * n) + 15) / 16 * 16;
    A.cp = n == 2 ? MctsCLay<2>::CP : n == 3 ? MctsCLay<3>::CP : MctsCLay<4>::CP;
    A.max_depth = 62 * n + 8;
    const size_t pool_bytes = (size_t)cap * (32 + A.cp) + (size_t)ecap * 24;
    A.n_pool_pages = (uint32_t)(pool_bytes / (MCTS_PAGE_UNITS * MCTS_UNIT)) + 4;
    A.max_pages = 2 * (int)A.n_pool_pages;
    A.pool = (uint8_t*)aligned_alloc(256, (size_t)A.n_pool_pages * MCTS_PAGE_UNITS * MCTS_UNIT);
    A.fq_slots = (uint32_t*)calloc(A.n_pool_pages, 4);
    A.fq_ctl = (int32_t*)calloc(4, 4);
    for (uint32_t p = 1; p < A.n_pool_pages; p++) A.fq_slots[p - 1] = p;
    A.fq_ctl[0] = 0; A.fq_ctl[1] = (int32_t)A.n_pool_pages - 1; A.fq_ctl[2] = (int32_t)A.n_pool_pages - 1; A.fq_ctl[3] = A.fq_ctl[2];
    A.tree_pages = (uint32_t*)calloc(A.max_pages, 4);
    A.htab = (uint32_t*)aligned_alloc(16, (size_t)A.hcap * 4);
    memset(A.htab, 0, (size_t)A.hcap * 4);
    A.trees = (MctsTree*)calloc(1, sizeof(MctsTree));
    A.path = (uint32_t*)calloc((size_t)A.max_depth * 2, 4);
    A.leaf_src = (uint8_t*)calloc(1, 1);
    for (int j = 0; j < MCTS_KMAX; j++) A.trees[0].slot[j].pend_edge = -1;
    m->P.cpuct = cpuct; m->P.fpu = fpu; m->P.temperature0 = temperature0; m->P.dirichlet_alpha = 0.3; m->P.seed = 0; m->P.game_base = 0;
    m->P.rules.limit = limit; m->P.rules.flags = rule_flags;
    m->edge_reserve = edge_reserve; m->gc_reachable = gc_reachable;
    m->leaf_state = (int8_t*)calloc(A.sp, 1);
    m->leaf_valid = (uint8_t*)calloc(SPL_ACTIONS, 1);
    m->pi = (float*)calloc(SPL_ACTIONS, 4);
    m->v = (float*)calloc(4, 4);
    return m;
}
void hm_destroy(HsMcts* m) {
    free(m->A.pool); free(m->A.fq_slots); free(m->A.fq_ctl); free(m->A.tree_pages); free(m->A.htab); free(m->A.trees); free(m->A.path); free(m->A.leaf_src);
    free(m->leaf_state); free(m->leaf_valid); free(m->pi); free(m->v); free(m);
}
int hm_free_pages(HsMcts* m) { return m->A.fq_ctl[2]; }
int hm_total_pages(HsMcts* m) { return (int)m->A.n_pool_pages - 1; }
void hm_reset(HsMcts* m) {
    MctsWarp w{0};
    mcts_clear_tree(w, m->A, 0);
    m->A.trees[0].nn_calls = 0;
}
// one getActionProb: begin + (descend, [rules, attach], fixed network, expand) until the budget is spent - the same
// sequence of steps the wave kernels run. Returns the tree status bits.
static void hm_maybe_clean(HsMcts* m) {
    if (m->clean_every <= 0) return;
    if (++m->clean_counter % m->clean_every) return;
    MctsWarp w{0};
    mcts_clean_tree(w, m->A, 0, 0, m->clean_gc);      // threshold 0: always compacts
}
void hm_set_episode(HsMcts* m, uint32_t e) { m->episode = e; }
void hm_set_clean(HsMcts* m, int every, int gc_reachable) { m->clean_every = every; m->clean_gc = gc_reachable; m->clean_counter = 0; }
int hm_search(HsMcts* m, const int8_t* root, int sims, uint32_t flags, const double* dir) {
    MctsWarp w{0};
    alignas(16) int8_t st[640];
    uint32_t scratch[24];
    double dscratch[4];
    alignas(16) uint8_t cst[128];
    DISPATCH(m->n, mcts_begin_tree<N>(w, m->A, 0, m->P, root, sims, flags, m->gc_reachable, dir, m->episode, st, cst, scratch, dscratch));
    for (;;) {
        int r = 0;
        DISPATCH(m->n, (r = mcts_descend_tree<N, false>(w, m->A, 0, 0, m->P, 2, 3, m->leaf_state, m->leaf_valid)));
        hm_maybe_clean(m);
        if (r == 0) break;
        if (r == 3) continue;
        if (r == 2) {
            bool ended = false; float es[4] = {0, 0, 0, 0}; uint32_t mask[13];
            memset(st, 0, sizeof st);
            DISPATCH(m->n, hm_step_rules<N>(m, st, &ended, es, mask));
            DISPATCH(m->n, (r = mcts_attach_tree<N, false>(w, m->A, 0, 0, m->P, st, cst, ended, es, mask, 1, m->leaf_state, m->leaf_valid)));
            hm_maybe_clean(m);
            if (r == 0) continue;
        }
        DISPATCH(m->n, mcts_fixed_net_row<N>(w, m->leaf_state, m->leaf_valid, m->pi, m->v, scratch));
        DISPATCH(m->n, (mcts_expand_tree<N, false>(w, m->A, 0, 0, m->P, m->pi, m->v, dir, dscratch)));
    }
    return (int)m->A.trees[0].status;
}
void hm_policy(HsMcts* m, double temp, double* probs, double* q) {
    MctsWarp w{0};
    double dscratch[4];
    DISPATCH(m->n, mcts_policy_tree<N>(w, m->A, 0, temp, probs, q, dscratch));
}
void hm_root_stats(HsMcts* m, int32_t* nsa, double* qsa, float* ps, int32_t* info8) {
    MctsWarp w{0};
    mcts_root_stats_tree(w, m->A, 0, nsa, qsa, ps, info8);
}
int hm_codec_roundtrip(int n, const int8_t* aos, int8_t* back, uint8_t* cst_out) {   // encode -> decode; returns 1 if the encoder accepted the state
    MctsWarp w{0};
    alignas(16) uint8_t cst[128];
    memset(cst, 0, sizeof cst);
    bool ok = false;
    DISPATCH(n, (ok = mcts_encode<N, true>(w, aos, cst)));
    if (ok) { DISPATCH(n, mcts_decode<N>(w, cst, back)); }
    memcpy(cst_out, cst, 128);
    return ok ? 1 : 0;
}
void hm_fixed_net(int n, const int8_t* state, const uint8_t* valid, float* pi, float* v) {
    MctsWarp w{0};
    uint32_t scratch[16];
    DISPATCH(n, mcts_fixed_net_row<N>(w, state, valid, pi, v, scratch));
}
}
