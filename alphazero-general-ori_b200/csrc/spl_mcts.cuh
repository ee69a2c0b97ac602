// GPU tree arena for MCTS - per-tree logic (device code; also compiled by g++ with a one-lane "warp"
// for the test-only host simulator in tests/hostsim/, never by the product's host path).
//
// What it replaces in the reference (MCTS.py):
//   mcts_descend_tree  <- MCTS.search :99-177 down to the first edge that was never taken, incl. pick_highest_UCB :199-219
//   mcts_rules_core    <- get_next_best_action_and_canonical_state :222-237 (make_move deterministic + swap_players) and
//                         the new node's getGameEnded :124 / getValidMoves :136 - run one LANE per tree (32 trees per warp)
//   mcts_attach_tree   <- the dictionary lookup / insertion of the child (nodes_data.get :120, :146) and leaf hand-over
//   mcts_expand_tree   <- the "first time that we explore state s" branch :134-148 (Ps from the network, normalise,
//                         root softmax + Dirichlet noise :141-144,180-186) followed by the backup :168-177
//   mcts_begin_tree    <- the dictionary lookup of the root (nodes_data.get :119-120) + tree cleaning :80-85
//   mcts_policy_tree   <- getActionProb's tail :61-97 (counts, policy-target pruning, temperature)
//
// One tree = one game lane = one warp. The reference's `nodes_data` dictionary (exact state bytes -> node, a DAG with
// transpositions that persists across moves) becomes a per-tree node pool + open-addressing hash table keyed by a
// 64-bit state hash with a full-state compare. Nodes are either TERMINAL (Es stored), NEEDS_NN (legal edges allocated,
// waiting for the network) or EXPANDED. Edges are sparse (only legal actions, in action order) and cache the child node
// index, so an inner traversal touches no rules code at all.
//
// Numerics follow the reference exactly: Ps float32, Qsa float64 running mean, Qs float32, Nsa/Ns integers, the -42
// "unvisited" sentinel, strict > in the arg-max (lowest action wins ties), forced-playout early return. All
// floating-point statements use explicitly rounded single operations (no FMA contraction).
#pragma once
#include <math.h>
#include "spl_rules.cuh"

#define MCTS_UNVISITED (-42.0)   // NAN sentinel, MCTS.py:9
#define MCTS_EPS 1e-8            // :8
#define MCTS_KFORCED 0.5         // :10

#ifdef __CUDACC__
#define MC_DMUL(a, b) __dmul_rn((a), (b))
#define MC_DADD(a, b) __dadd_rn((a), (b))
#define MC_DDIV(a, b) __ddiv_rn((a), (b))
#define MC_DSQRT(a) __dsqrt_rn((a))
#define MC_FMUL(a, b) __fmul_rn((a), (b))
#define MC_FADD(a, b) __fadd_rn((a), (b))
#define MC_FDIV(a, b) __fdiv_rn((a), (b))
#else   // host simulator: compiled with -ffp-contract=off
#define MC_DMUL(a, b) ((a) * (b))
#define MC_DADD(a, b) ((a) + (b))
#define MC_DDIV(a, b) ((a) / (b))
#define MC_DSQRT(a) sqrt((a))
#define MC_FMUL(a, b) ((a) * (b))
#define MC_FADD(a, b) ((a) + (b))
#define MC_FDIV(a, b) ((a) / (b))
#endif

// ------------------------------------------------------------------------------------------
// warp policy: 32 lanes on the device, 1 lane in the host simulator
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
struct MctsWarp {
    int lane;
    static constexpr int W = 32;
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ uint32_t lanemask_lt() const { return (1u << lane) - 1u; }
    __device__ __forceinline__ int sum(int v) const { return __reduce_add_sync(0xffffffffu, v); }
    __device__ __forceinline__ uint64_t sum64(uint64_t v) const {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ double sumd_tree(double v) const {   // fixed butterfly order (production-only sums)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v = MC_DADD(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ __forceinline__ int shfl(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    // arg-max, lowest index wins ties; idx < 0 = none. Three warp reductions (REDUX) on an order-preserving integer image of the
    // double instead of five shuffle rounds: max of the high words, max of the low words among those, min index among those.
    // u + 0.0 maps -0.0 to +0.0 so that equal doubles have equal images (no NaNs occur).
    __device__ __forceinline__ void best(double& u, int& idx) const {
        const long long b = __double_as_longlong(u + 0.0);
        const unsigned long long key = idx < 0 ? 0ull : (b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull));
        const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
        const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
        const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        const bool win = idx >= 0 && hi == mh && lo == ml;
        const int bi = __reduce_min_sync(0xffffffffu, win ? idx : 0x7fffffff);
        idx = bi == 0x7fffffff ? -1 : bi;
    }
    // the two square roots of pick_highest_UCB in ONE pass through the sqrt sequence: odd lanes take b, even lanes a
    __device__ __forceinline__ void sqrt2(double a, double b, double& ra, double& rb) const {
        const double r = MC_DSQRT((lane & 1) ? b : a);
        ra = __shfl_sync(0xffffffffu, r, 0);
        rb = __shfl_sync(0xffffffffu, r, 1);
    }
};
#else
struct MctsWarp {
    int lane;
    static constexpr int W = 1;
    void sync() const {}
    uint32_t ballot(bool p) const { return p ? 1u : 0u; }
    uint32_t lanemask_lt() const { return 0u; }
    int sum(int v) const { return v; }
    uint64_t sum64(uint64_t v) const { return v; }
    double sumd_tree(double v) const { return v; }
    int shfl(int v, int) const { return v; }
    void best(double&, int&) const {}
    void sqrt2(double a, double b, double& ra, double& rb) const { ra = MC_DSQRT(a); rb = MC_DSQRT(b); }
};
#endif

// ------------------------------------------------------------------------------------------
// arena layout (all in caller-owned HBM)
// ------------------------------------------------------------------------------------------
enum { MCTS_NODE_TERMINAL = 1, MCTS_NODE_NEEDS_NN = 2, MCTS_NODE_EXPANDED = 3 };
enum { MCTS_F_FORCED = 1u, MCTS_F_NOISE = 2u };                       // per-move flags (getActionProb :54-58)
enum { MCTS_S_OVERFLOW_NODES = 1u, MCTS_S_OVERFLOW_EDGES = 2u, MCTS_S_PROTOCOL = 4u };   // sticky status bits

struct MctsNode {   // 32 B
    uint64_t hash;
    uint8_t ply, kind;
    uint16_t n_edges;
    uint32_t edge_off;
    union {
        struct { int32_t Ns; float Qs; uint32_t pad[2]; } x;   // EXPANDED / NEEDS_NN
        float es[4];                                          // TERMINAL: getGameEnded vector
    } u;
};
struct MctsEdge {   // 32 B
    double Q;        // Qsa (float64, -42 = unvisited)
    float P;         // Ps[a]
    int32_t N;       // Nsa
    uint32_t child;  // node index + 1 (0 = not linked yet)
    uint16_t action;
    uint16_t child_ne;    // the child's edge count and first edge, copied here when the edge is linked: the descent then
    uint32_t child_eoff;  // fetches the child's header and its edges in ONE round trip instead of two dependent ones
    uint32_t pad;
};
struct MctsTree {   // 96 B
    int32_t n_nodes, n_edges, root, leaf;
    int32_t sims_done, sims_target, path_len;
    uint32_t flags, status;
    int32_t nn_calls, resets, compactions;
    float last_v[4];   // value vector the last finished simulation returned at the root (what MCTS.search returns)
    int32_t truncated; // searches cut short because a pool filled up mid-move (the next begin makes room again)
    int32_t depth_sum; // sum of path lengths of the simulations since the last reset (diagnostics)
    // state of the simulation in flight (it survives between the kernels of a wave):
    int32_t cur;         // >= 0: continue the descent at this node with path_len edges already recorded; -1: start at the root
    int32_t pend_edge;   // >= 0: the descent stopped at this (absolute) edge, whose child state is being computed / attached
    int32_t pend_parent; // node index of that edge's parent
    int32_t spec_hits;   // diagnostics: levels of the descents that started from the early-fetched child (see mcts_descend_tree)
    int32_t pad[2];
};
struct MctsArena {
    int n_trees, cap, ecap, hcap, sp, max_depth;
    int8_t* states;    // [T][cap][sp]   node states, the reference's int8[R,7] bytes + zero padding to 16
    MctsNode* nodes;   // [T][cap]
    MctsEdge* edges;   // [T][ecap]
    uint32_t* htab;    // [T][hcap]      node index + 1, linear probing
    MctsTree* trees;   // [T]
    uint32_t* path;    // [T][max_depth][2]  (node, absolute edge index) of the current simulation
    // staging between the rules kernel and the attach kernel (the child state of every tree's pending edge)
    int8_t* stage_state;   // [T][sp]
    uint32_t* stage_mask;  // [13][T]
    float* stage_es;       // [T][4]
    uint8_t* stage_ended;  // [T]
    long long* prof;       // diagnostics (normally NULL): [T][16] per-tree time stamps of the last wave (spl_mcts_debug_profile)
    uint8_t* leaf_src;     // [T]  where the network input of the tree's leaf lives: 0 = its leaf row (bytes), 1 = the staging row
                           //      (state bytes + mask words) the rules kernel wrote - lets the network start before the attach
};
struct MctsSearchParams {
    double cpuct, fpu, temperature0, dirichlet_alpha;
    uint64_t seed;
    uint32_t game_base;
    SplRules rules;
};

template <int N> struct MctsLay {
    static constexpr int S = SplLay<N>::CELLS;
    static constexpr int SP = (S + 15) / 16 * 16;
    static constexpr int MAX_DEPTH = 62 * N + 8;
};

#ifdef __CUDACC__
#define SPL_M __device__ __forceinline__
#else
#define SPL_M inline
#endif
struct AosAcc {   // the reference's own array order: cell (row, col) = byte 7*row + col
    int8_t* p;
    SPL_M int get(int row, int col) const { return p[7 * row + col]; }
    SPL_M void set(int row, int col, int v) { p[7 * row + col] = (int8_t)v; }
};

SPL_D uint64_t mcts_mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    return x;
}

// 64-bit hash of a padded state (sp bytes, 8-byte aligned): a sum of position-salted word mixes, so that any
// partition of the words over lanes gives the same value
template <class W>
SPL_D uint64_t mcts_hash(const W& w, const int8_t* st, int sp) {
    const uint64_t* q = reinterpret_cast<const uint64_t*>(st);
    uint64_t h = 0;
    for (int i = w.lane; i < sp / 8; i += W::W) h += mcts_mix64(q[i] + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull);
    return mcts_mix64(w.sum64(h));
}

template <class W>
SPL_D void mcts_copy16(const W& w, void* dst, const void* src, int bytes) {   // bytes % 16 == 0, both 16-aligned
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = w.lane; i < bytes / 16; i += W::W) d[i] = s[i];
}

template <class W>
SPL_D bool mcts_equal16(const W& w, const void* a, const void* b, int bytes) {
    const uint4* x = reinterpret_cast<const uint4*>(a);
    const uint4* y = reinterpret_cast<const uint4*>(b);
    bool diff = false;
    for (int i = w.lane; i < bytes / 16; i += W::W) {
        const uint4 p = x[i], q = y[i];
        diff |= (p.x != q.x) | (p.y != q.y) | (p.z != q.z) | (p.w != q.w);
    }
    return w.ballot(diff) == 0u;
}

// nodes_data.get(s) :120
template <class W>
SPL_D int mcts_lookup(const W& w, const MctsArena& A, int t, const int8_t* st, uint64_t h) {
    const uint32_t* tab = A.htab + (size_t)t * A.hcap;
    const MctsNode* nodes = A.nodes + (size_t)t * A.cap;
    uint32_t slot = (uint32_t)h & (uint32_t)(A.hcap - 1);
    for (;;) {
        const uint32_t e = tab[slot];
        if (e == 0u) return -1;
        const int idx = (int)e - 1;
        if (nodes[idx].hash == h && mcts_equal16(w, A.states + ((size_t)t * A.cap + idx) * A.sp, st, A.sp)) return idx;
        slot = (slot + 1u) & (uint32_t)(A.hcap - 1);
    }
}

template <class W>
SPL_D void mcts_table_insert(const W& w, const MctsArena& A, int t, int idx, uint64_t h) {
    if (w.lane == 0) {
        uint32_t* tab = A.htab + (size_t)t * A.hcap;
        uint32_t slot = (uint32_t)h & (uint32_t)(A.hcap - 1);
        while (tab[slot] != 0u) slot = (slot + 1u) & (uint32_t)(A.hcap - 1);
        tab[slot] = (uint32_t)idx + 1u;
    }
    w.sync();
}

// New node for the state `st` (sp bytes, zero padded) whose end-of-game vector / legal mask are already known:
// stores the bytes, allocates one edge per legal action (in action order), links it into the hash table.
// m: the 13 mask words at m[i * mstride]; es: N floats (used when ended). Returns the node index or -1 on pool overflow.
template <int N, class W>
SPL_D int mcts_store_node(const W& w, const MctsArena& A, int t, const int8_t* st, uint64_t h, bool ended, const float* es,
                          const uint32_t* m, int mstride) {
    MctsTree* T = A.trees + t;
    const int idx = T->n_nodes;
    if (idx >= A.cap) {
        if (w.lane == 0) T->status |= MCTS_S_OVERFLOW_NODES;
        w.sync();
        return -1;
    }
    int k = 0;
    if (!ended)
        for (int i = 0; i < SPL_MASK_WORDS; i++) k += SPL_POPC(m[i * mstride]);
    const int e0 = T->n_edges;
    if (!ended && e0 + k > A.ecap) {
        if (w.lane == 0) T->status |= MCTS_S_OVERFLOW_EDGES;
        w.sync();
        return -1;
    }
    MctsNode* nd = A.nodes + (size_t)t * A.cap + idx;
    if (w.lane == 0) {
        nd->hash = h;
        nd->ply = (uint8_t)st[6];
        nd->n_edges = (uint16_t)k;
        nd->edge_off = (uint32_t)e0;
        if (ended) {
            nd->kind = MCTS_NODE_TERMINAL;
            for (int i = 0; i < 4; i++) nd->u.es[i] = i < N ? es[i] : 0.f;
        } else {
            nd->kind = MCTS_NODE_NEEDS_NN;
            nd->u.x.Ns = 0; nd->u.x.Qs = 0.f; nd->u.x.pad[0] = nd->u.x.pad[1] = 0u;
        }
    }
    mcts_copy16(w, A.states + ((size_t)t * A.cap + idx) * A.sp, st, A.sp);
    if (!ended) {   // edges in action order: word i of the mask owns a contiguous run
        MctsEdge* ed = A.edges + (size_t)t * A.ecap + e0;
        for (int i = w.lane; i < SPL_MASK_WORDS; i += W::W) {
            int off = 0;
            for (int j = 0; j < i; j++) off += SPL_POPC(m[j * mstride]);
            uint32_t bits = m[i * mstride];
            while (bits) {
                const int b = SPL_FFS(bits) - 1;
                bits &= bits - 1u;
                MctsEdge e;
                e.Q = MCTS_UNVISITED; e.P = 0.f; e.N = 0; e.child = 0u; e.action = (uint16_t)(32 * i + b); e.child_ne = 0; e.child_eoff = 0u; e.pad = 0u;
                ed[off++] = e;
            }
        }
    }
    w.sync();
    if (w.lane == 0) {
        T->n_nodes = idx + 1;
        if (!ended) T->n_edges = e0 + k;
    }
    mcts_table_insert(w, A, t, idx, h);
    return idx;
}

// getGameEnded (:124) and, for a live position, getValidMoves (:136) of a canonical state, as one thread's work
template <int N, class S>
SPL_D bool mcts_eval_state(const S& s, SplRules rules, float* es, uint32_t* m) {
    const bool ended = spl_game_ended<N>(s, rules, es);
    if (!ended) spl_valid_mask<N>(s, 0, rules, m);
    else
        for (int i = 0; i < SPL_MASK_WORDS; i++) m[i] = 0u;
    return ended;
}

// in-tree step of one thread: make_move(a, 0, deterministic=True) + swap_players(next_player) (:226-235), then the
// evaluation of the new state. One LANE per tree on the device (32 trees per warp on shared-memory tiles).
template <int N, class S>
SPL_D bool mcts_rules_core(S& s, int action, SplRules rules, float* es, uint32_t* m) {
    SplChance ch;
    ch.mode = 0; ch.code = 0; ch.seed = 0; ch.game = 0; ch.episode = 0; ch.ply = 0;
    const int nxt = spl_apply_move<N>(s, action, 0, ch);
    if (nxt > 0) spl_rotate<N>(s, nxt, rules);
    return mcts_eval_state<N>(s, rules, es, m);
}

// root creation at the start of a move (one per move per tree: lane 0 evaluates the state, the warp stores it)
// `scratch` = 24 uint32 of per-warp scratch.
template <int N, class W>
SPL_D int mcts_create_node(const W& w, const MctsArena& A, int t, const MctsSearchParams& P, int8_t* st, uint64_t h, uint32_t* scratch) {
    if (w.lane == 0) {
        AosAcc s{st};
        float es[N];
        const bool ended = mcts_eval_state<N>(s, P.rules, es, scratch);
        scratch[13] = ended ? 1u : 0u;
        for (int i = 0; i < N; i++) memcpy(&scratch[16 + i], &es[i], 4);
    }
    w.sync();
    float es[N];
    for (int i = 0; i < N; i++) memcpy(&es[i], &scratch[16 + i], 4);
    const int idx = mcts_store_node<N>(w, A, t, st, h, scratch[13] != 0u, es, scratch, 1);
    w.sync();
    return idx;
}

// ------------------------------------------------------------------------------------------
// root softmax + Dirichlet noise (softmax :244-250, applyDirNoise :180-186, normalise :239-242)
// ------------------------------------------------------------------------------------------
// Marsaglia-Tsang gamma(alpha,1) for the on-device Dirichlet sampler (production; parity runs inject the vector)
SPL_D double mcts_u01(uint32_t a, uint32_t b) { return ((double)(((uint64_t)a << 21) ^ (uint64_t)(b >> 11)) + 0.5) * (1.0 / 9007199254740992.0); }
SPL_D double mcts_gamma(double alpha, uint64_t seed, uint32_t game, uint32_t ply, uint32_t k) {
    const double a1 = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = a1 - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double g = d;
    for (uint32_t attempt = 0; attempt < 64u; attempt++) {
        const SplPhilox r0 = spl_philox(seed, game, ply, k * 128u + 2u * attempt, 4);
        const SplPhilox r1 = spl_philox(seed, game, ply, k * 128u + 2u * attempt + 1u, 4);
        const double u1 = mcts_u01(r0.v[0], r0.v[1]), u2 = mcts_u01(r0.v[2], r0.v[3]), u3 = mcts_u01(r1.v[0], r1.v[1]);
        const double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(u3) < 0.5 * x * x + d - d * v + d * log(v)) {
            g = d * v;
            if (alpha < 1.0) g *= pow(mcts_u01(r1.v[2], r1.v[3]), 1.0 / alpha);
            break;
        }
    }
    return g;
}

// Ps <- normalise(0.75 * softmax(Ps, T0) + 0.25 * Dir) on the node's edges. dir (may be NULL): the vector
// rng.dirichlet returned, one value per legal action in action order.
template <class W>
SPL_D void mcts_root_noise(const W& w, MctsEdge* ed, int k, const MctsSearchParams& P, const double* dir, uint32_t game, uint32_t ply,
                           double* dscratch /* >= 2 doubles per warp */) {
    if (P.temperature0 != 1.0) {   // Ps ** (1/T) in float64, normalised, back to float32 (softmax :247-250)
        const double inv_t = 1.0 / P.temperature0;
        if (w.lane == 0) {
            double s = 0.0;
            for (int i = 0; i < k; i++) s = MC_DADD(s, pow((double)ed[i].P, inv_t));
            dscratch[0] = s;
        }
        w.sync();
        const double s = dscratch[0];
        for (int i = w.lane; i < k; i += W::W) ed[i].P = (float)MC_DDIV(pow((double)ed[i].P, inv_t), s);
        w.sync();
    }
    double gsum = 1.0;
    if (!dir) {   // Dirichlet(alpha) = iid gamma(alpha) / their sum; the gammas are recomputed below (counter-based)
        double part = 0.0;
        for (int i = w.lane; i < k; i += W::W) part = MC_DADD(part, mcts_gamma(P.dirichlet_alpha, P.seed, game, ply, (uint32_t)i));
        gsum = w.sumd_tree(part);
    }
    for (int i = w.lane; i < k; i += W::W) {
        const double dv = dir ? dir[i] : MC_DDIV(mcts_gamma(P.dirichlet_alpha, P.seed, game, ply, (uint32_t)i), gsum);
        ed[i].P = (float)MC_DADD((double)MC_FMUL(0.75f, ed[i].P), MC_DMUL(0.25, dv));
    }
    w.sync();
    if (w.lane == 0) {   // np.sum in float32, action order
        float s = 0.f;
        for (int i = 0; i < k; i++) s = MC_FADD(s, ed[i].P);
        reinterpret_cast<float*>(dscratch)[0] = s;
    }
    w.sync();
    const float s = reinterpret_cast<float*>(dscratch)[0];
    for (int i = w.lane; i < k; i += W::W) ed[i].P = MC_FDIV(ed[i].P, s);
    w.sync();
}

// ------------------------------------------------------------------------------------------
// pick_highest_UCB :199-219 over the node's edges (already restricted to legal actions, in action order)
// returns the edge position inside the node
// ------------------------------------------------------------------------------------------
template <class W>
SPL_D int mcts_pick(const W& w, const MctsEdge* ed, const MctsEdge& first, int k, int Ns, float Qs, const MctsSearchParams& P, bool forced,
                    int n_iter) {   // `first` = ed[lane], fetched by the caller together with the node header
    const double fpu_init = P.fpu > 0.0 ? MC_DADD((double)Qs, -P.fpu) : P.fpu;   // :202
    double sq_ns, sq_ns_eps;
    w.sqrt2((double)Ns, MC_DADD((double)Ns, MCTS_EPS), sq_ns, sq_ns_eps);
    double best_u = 0.0;
    int best_i = -1, forced_i = 0x7fffffff;
    for (int i = w.lane; i < k; i += W::W) {
        const MctsEdge e = i == w.lane ? first : ed[i];
        if (forced && forced_i == 0x7fffffff) {   // :207-208 - the first legal action short of its forced visits wins outright
            const long long quota = (long long)MC_DSQRT(MC_DMUL(MC_DMUL(MCTS_KFORCED, (double)e.P), (double)n_iter));
            if ((long long)e.N < quota) forced_i = i;
        }
        double u;
        if (e.Q != MCTS_UNVISITED) u = MC_DADD(e.Q, MC_DDIV(MC_DMUL(MC_DMUL(P.cpuct, (double)e.P), sq_ns), (double)(1 + e.N)));   // :211
        else u = MC_DADD(fpu_init, MC_DMUL(MC_DMUL(P.cpuct, (double)e.P), sq_ns_eps));                                                 // :213
        if (best_i < 0 ? (u > -INFINITY) : (u > best_u)) { best_u = u; best_i = i; }                                                   // :215 strict >
    }
    if (forced) {
        int f = forced_i;
#ifdef __CUDACC__
        f = __reduce_min_sync(0xffffffffu, f);
#endif
        if (f != 0x7fffffff) return f;
    }
    w.best(best_u, best_i);
    return best_i;
}

// backup along the recorded path (:168-177). v = value vector in the frame of the node below the last edge (the same in
// every lane). np.roll(v, next_player) with next_player = 1 after every in-tree move, so the edge at depth d sees the
// vector rolled (depth - d) times and uses its component 0 = v[(d - depth) mod N]. A path never holds a node or an edge
// twice (the ply grows along it), so the levels are independent: one lane per level.
template <int N>
SPL_D float mcts_vsel(const float* v, int i) {
    float r = v[0];
#pragma unroll
    for (int j = 1; j < N; j++) r = i == j ? v[j] : r;
    return r;
}
template <int N, class W>
SPL_D void mcts_backup(const W& w, const MctsArena& A, int t, int depth, const float* v) {
    MctsNode* nodes = A.nodes + (size_t)t * A.cap;
    MctsEdge* edges = A.edges + (size_t)t * A.ecap;
    const uint32_t* path = A.path + (size_t)t * A.max_depth * 2;
    for (int d = w.lane; d < depth; d += W::W) {
        const float v0 = mcts_vsel<N>(v, ((d - depth) % N + N) % N);
        MctsNode* nd = nodes + path[2 * d];
        MctsEdge* e = edges + path[2 * d + 1];
        e->Q = MC_DDIV(MC_DADD(MC_DMUL((double)e->N, e->Q), (double)v0), (double)(e->N + 1));                               // :171
        nd->u.x.Qs = MC_FDIV(MC_FADD(MC_FMUL((float)(nd->u.x.Ns + 1), nd->u.x.Qs), v0), (float)(nd->u.x.Ns + 2));          // :172
        e->N += 1;
        nd->u.x.Ns += 1;
    }
    if (w.lane == 0) {
        MctsTree* T = A.trees + t;
#pragma unroll
        for (int i = 0; i < 4; i++) T->last_v[i] = i < N ? mcts_vsel<N>(v, ((i - depth) % N + N) % N) : 0.f;
        T->depth_sum += depth;
    }
}

// hands a node that waits for the network to the leaf row of its tree (:136-138)
template <int N, class W>
SPL_D void mcts_emit_leaf(const W& w, const MctsArena& A, int t, int node, int depth, int sims_done, int8_t* leaf_state, uint8_t* leaf_valid,
                          bool rows = true) {   // rows = false: the staging row of the rules kernel already holds this very state and mask
    typedef MctsLay<N> ML;
    MctsTree* T = A.trees + t;
    if (rows) {
        const MctsNode* nd = A.nodes + (size_t)t * A.cap + node;
        const int8_t* src = A.states + ((size_t)t * A.cap + node) * A.sp;
        for (int i = w.lane; i < ML::S; i += W::W) leaf_state[i] = src[i];
        for (int i = w.lane; i < SPL_ACTIONS; i += W::W) leaf_valid[i] = 0;
        w.sync();
        const MctsEdge* ed = A.edges + (size_t)t * A.ecap + nd->edge_off;
        for (int i = w.lane; i < (int)nd->n_edges; i += W::W) leaf_valid[ed[i].action] = 1;
        if (w.lane == 0) A.leaf_src[t] = 0;
    }
    if (w.lane == 0) { T->leaf = node; T->path_len = depth; T->sims_done = sims_done; T->cur = -1; T->pend_edge = -1; }
    w.sync();
}

// ------------------------------------------------------------------------------------------
// descent (light: no rules code): runs simulations of tree t until one
//   * reaches a node that waits for the network  -> leaf row written, returns 1
//   * reaches an edge that was never taken       -> records it as pending (the rules kernel computes the child state,
//                                                    mcts_attach_tree links it), returns 2
//   * or the move's budget is spent              -> returns 0
// Simulations that end in a terminal node are backed up on the spot, at most `max_terminal` of them per call (then 3 is
// returned and the next call carries on): near the end of a game almost every simulation is such a one, and a single
// tree working through its whole budget would hold up the wave. For the same reason a call walks at most `max_levels`
// edges before it yields (3). Trees with a leaf or a pending edge are skipped.
// ------------------------------------------------------------------------------------------
template <int N, class W>
SPL_D int mcts_descend_tree(const W& w, const MctsArena& A, int t, const MctsSearchParams& P, int max_terminal, int max_levels,
                            int8_t* leaf_state, uint8_t* leaf_valid) {
    MctsTree* T = A.trees + t;
    if (T->leaf >= 0) return 1;
    if (T->pend_edge >= 0) return 2;
    if (T->status != 0u) return 0;
    MctsNode* nodes = A.nodes + (size_t)t * A.cap;
    MctsEdge* edges = A.edges + (size_t)t * A.ecap;
    uint32_t* path = A.path + (size_t)t * A.max_depth * 2;
    int sims_done = T->sims_done;
    const int target = T->sims_target;
    const uint32_t flags = T->flags;
    const int root = T->root;
    int cur = T->cur, depth = cur >= 0 ? T->path_len : 0;
    if (cur < 0) cur = root;
    // where the edges of `cur` are expected: known from the parent's edge, or from the node header at the start of a walk
    uint32_t eoff = nodes[cur].edge_off;
    int ne = nodes[cur].kind == MCTS_NODE_TERMINAL ? 0 : (int)nodes[cur].n_edges;
    // The child the PREVIOUS visit of a node chose is fetched (header + edges) while this visit is still computing its pick: the
    // search tends to walk the same line again, and then the next level starts without a memory round trip. The hint lives in
    // a spare word of the node header (edge position + 1); it never influences a result, only what is loaded early.
    bool have_spec = false;
    MctsEdge spec_first;
    MctsNode spec_nd;
    while (sims_done < target) {
        for (;;) {
            const MctsEdge* ed = edges + eoff;
            MctsEdge first;                                  // header and first 32 edges: one round trip
            MctsNode nd;
            if (have_spec) {
                first = spec_first; nd = spec_nd;
            } else {
                first.Q = MCTS_UNVISITED; first.P = 0.f; first.N = 0; first.child = 0u; first.action = 0; first.child_ne = 0; first.child_eoff = 0u; first.pad = 0u;
                if (w.lane < ne) first = ed[w.lane];
                nd = nodes[cur];
            }
            have_spec = false;
            const int kind = nd.kind;
            if (kind == MCTS_NODE_NEEDS_NN) {
                mcts_emit_leaf<N>(w, A, t, cur, depth, sims_done, leaf_state, leaf_valid);
                return 1;
            }
            if (kind == MCTS_NODE_TERMINAL) break;   // :130-132
            int hp = -1;
            const int hint = (int)nd.u.x.pad[0] - 1;
#ifdef __CUDACC__
            if (hint >= 0 && hint < 32 && hint < (int)nd.n_edges) {
                const uint32_t hchild = __shfl_sync(0xffffffffu, first.child, hint);
                const uint32_t heoff = __shfl_sync(0xffffffffu, first.child_eoff, hint);
                const int hne = __shfl_sync(0xffffffffu, (int)first.child_ne, hint);
                if (hchild != 0u) {
                    hp = hint;
                    spec_first.Q = MCTS_UNVISITED; spec_first.P = 0.f; spec_first.N = 0; spec_first.child = 0u; spec_first.action = 0;
                    spec_first.child_ne = 0; spec_first.child_eoff = 0u; spec_first.pad = 0u;
                    if (w.lane < hne) spec_first = edges[heoff + w.lane];
                    spec_nd = nodes[hchild - 1u];
                }
            }
#endif
#ifdef __CUDACC__
            long long prof_t0 = 0;
            if (A.prof) prof_t0 = clock64();
#endif
            const int ei = mcts_pick(w, ed, first, (int)nd.n_edges, nd.u.x.Ns, nd.u.x.Qs, P, depth == 0 && (flags & MCTS_F_FORCED), sims_done);
#ifdef __CUDACC__
            if (A.prof && w.lane == 0) { A.prof[(size_t)t * 16 + 6] += clock64() - prof_t0; A.prof[(size_t)t * 16 + 7] += 1; }   // diagnostics: cycles inside the pick, levels
#endif
            if (w.lane == 0) {
                path[2 * depth] = (uint32_t)cur; path[2 * depth + 1] = nd.edge_off + (uint32_t)ei;
                if (ei != hint) nodes[cur].u.x.pad[0] = (uint32_t)ei + 1u;
            }
            depth++;
            uint32_t child, child_eoff;
            int child_ne;
#ifdef __CUDACC__
            if (ei < 32) {
                child = __shfl_sync(0xffffffffu, first.child, ei);
                child_eoff = __shfl_sync(0xffffffffu, first.child_eoff, ei);
                child_ne = __shfl_sync(0xffffffffu, (int)first.child_ne, ei);
            } else
#endif
            {
                const MctsEdge sel = ed[ei];
                child = sel.child; child_eoff = sel.child_eoff; child_ne = (int)sel.child_ne;
            }
            if (child != 0u && --max_levels <= 0) {   // yield: a very deep path finishes in the next call instead of holding up the wave
                if (w.lane == 0) { T->cur = (int)child - 1; T->path_len = depth; T->sims_done = sims_done; }
                w.sync();
                return 3;
            }
            if (child == 0u) {   // first traversal of this edge
                if (w.lane == 0) {
                    T->pend_edge = (int32_t)(nd.edge_off + (uint32_t)ei); T->pend_parent = cur;
                    T->path_len = depth; T->sims_done = sims_done; T->cur = -1;
                }
                w.sync();
                return 2;
            }
            cur = (int)child - 1;
            eoff = child_eoff; ne = child_ne;
            have_spec = ei == hp;
#ifdef __CUDACC__
            if (have_spec && w.lane == 0) atomicAdd(&T->spec_hits, 1);
#endif
        }
        {   // terminal: return Es up the path
            float v[N];
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = nodes[cur].u.es[i];
            mcts_backup<N>(w, A, t, depth, v);
        }
        w.sync();
        sims_done++;
        cur = root; depth = 0;
        eoff = nodes[root].edge_off; ne = nodes[root].kind == MCTS_NODE_TERMINAL ? 0 : (int)nodes[root].n_edges;
        if (--max_terminal <= 0 && sims_done < target) {
            if (w.lane == 0) { T->sims_done = sims_done; T->cur = -1; T->path_len = 0; }
            w.sync();
            return 3;
        }
    }
    if (w.lane == 0) { T->sims_done = sims_done; T->cur = -1; T->path_len = 0; }
    w.sync();
    return 0;
}

// attaches the child state computed for the pending edge: st = its bytes (sp, zero padded, 16-aligned), ended / es /
// mask from mcts_rules_core. The dictionary may already hold the state (transposition, :120): then the descent goes on
// from that node in the next mcts_descend_tree call. Returns 1 if a leaf row was written, 0 otherwise.
template <int N, class W>
SPL_D int mcts_attach_tree(const W& w, const MctsArena& A, int t, const MctsSearchParams& P, const int8_t* st, bool ended, const float* es,
                           const uint32_t* m, int mstride, int8_t* leaf_state, uint8_t* leaf_valid, bool emit_rows = true) {
    MctsTree* T = A.trees + t;
    const int pe = T->pend_edge;
    if (pe < 0) return T->leaf >= 0 ? 1 : 0;
    const uint64_t h = mcts_hash(w, st, A.sp);
    int idx = mcts_lookup(w, A, t, st, h);
    if (idx < 0) idx = mcts_store_node<N>(w, A, t, st, h, ended, es, m, mstride);
    if (idx < 0) {   // pool overflow: the search of this tree stops here (status bit set)
        if (w.lane == 0) { T->pend_edge = -1; T->cur = -1; T->path_len = 0; }
        w.sync();
        return 0;
    }
    MctsNode* nodes = A.nodes + (size_t)t * A.cap;
    if (w.lane == 0) {
        MctsEdge* e = A.edges + (size_t)t * A.ecap + pe;
        e->child = (uint32_t)idx + 1u;
        e->child_eoff = nodes[idx].edge_off;
        e->child_ne = nodes[idx].kind == MCTS_NODE_TERMINAL ? (uint16_t)0 : nodes[idx].n_edges;
    }
    w.sync();
    const int kind = nodes[idx].kind;
    const int depth = T->path_len, sims_done = T->sims_done;
    if (kind == MCTS_NODE_NEEDS_NN) {
        mcts_emit_leaf<N>(w, A, t, idx, depth, sims_done, leaf_state, leaf_valid, emit_rows);
        return 1;
    }
    if (kind == MCTS_NODE_TERMINAL) {
        float v[N];
#pragma unroll
        for (int i = 0; i < N; i++) v[i] = nodes[idx].u.es[i];
        mcts_backup<N>(w, A, t, depth, v);
        if (w.lane == 0) { T->sims_done = sims_done + 1; T->pend_edge = -1; T->cur = -1; T->path_len = 0; }
    } else {   // an expanded node reached through a new edge: keep descending from it
        if (w.lane == 0) { T->pend_edge = -1; T->cur = idx; }
    }
    w.sync();
    return 0;
}

// expansion + backup: pi = the network's probability row for the leaf (float32[406], masked softmax), v = float32[N]
template <int N, class W>
SPL_D void mcts_expand_tree(const W& w, const MctsArena& A, int t, const MctsSearchParams& P, const float* pi, const float* vin,
                            const double* dir, double* dscratch) {
    MctsTree* T = A.trees + t;
    const int leaf = T->leaf;
    if (leaf < 0) return;
    MctsNode* nd = A.nodes + (size_t)t * A.cap + leaf;
#ifdef __CUDACC__
    {
        // The common case in three rounds of loads instead of seven dependent ones: (1) the leaf's header, the value vector and
        // the recorded path, (2) the leaf's edge actions and the statistics of the path's nodes and edges, (3) the network's
        // probabilities; the float32 normalisation sums in action order through shuffles (same additions as `normalise :239`),
        // the backup (:168-177) uses the statistics fetched in round 2. Same arithmetic as the general path below.
        const int depth = T->path_len;
        const bool noise = leaf == T->root && T->sims_done == 0 && (T->flags & MCTS_F_NOISE);
        const MctsNode hdr = *nd;
        const int k = hdr.n_edges;
        if (!noise && k <= 32 && depth <= 32) {
            MctsNode* nodes = A.nodes + (size_t)t * A.cap;
            MctsEdge* edges = A.edges + (size_t)t * A.ecap;
            const uint32_t* path = A.path + (size_t)t * A.max_depth * 2;
            float v[N];
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = vin[i];
            uint32_t pn = 0u, pe = 0u;
            if (w.lane < depth) { pn = path[2 * w.lane]; pe = path[2 * w.lane + 1]; }
            MctsEdge* ed = edges + hdr.edge_off;
            int act = 0;
            if (w.lane < k) act = ed[w.lane].action;
            int eN = 0, nNs = 0;
            double eQ = 0.0;
            float nQs = 0.f;
            if (w.lane < depth) { eN = edges[pe].N; eQ = edges[pe].Q; nNs = nodes[pn].u.x.Ns; nQs = nodes[pn].u.x.Qs; }
            float p = w.lane < k ? pi[act] : 0.f;
            float s = 0.f;
            for (int i = 0; i < k; i++) s = MC_FADD(s, __shfl_sync(0xffffffffu, p, i));
            if (w.lane < k) ed[w.lane].P = MC_FDIV(p, s);
            if (w.lane == 0) {
                nd->u.x.Ns = 0;
                nd->u.x.Qs = v[0];   // :147
                nd->kind = MCTS_NODE_EXPANDED;
            }
            if (w.lane < depth) {
                const float v0 = mcts_vsel<N>(v, ((w.lane - depth) % N + N) % N);
                edges[pe].Q = MC_DDIV(MC_DADD(MC_DMUL((double)eN, eQ), (double)v0), (double)(eN + 1));                 // :171
                nodes[pn].u.x.Qs = MC_FDIV(MC_FADD(MC_FMUL((float)(nNs + 1), nQs), v0), (float)(nNs + 2));             // :172
                edges[pe].N = eN + 1;
                nodes[pn].u.x.Ns = nNs + 1;
            }
            if (w.lane == 0) {
#pragma unroll
                for (int i = 0; i < 4; i++) T->last_v[i] = i < N ? mcts_vsel<N>(v, ((i - depth) % N + N) % N) : 0.f;
                T->depth_sum += depth;
                T->sims_done += 1;
                T->leaf = -1; T->cur = -1; T->pend_edge = -1; T->path_len = 0;
                T->nn_calls += 1;
            }
            w.sync();
            return;
        }
    }
#endif
    MctsEdge* ed = A.edges + (size_t)t * A.ecap + nd->edge_off;
    const int k = nd->n_edges;
    for (int i = w.lane; i < k; i += W::W) ed[i].P = pi[ed[i].action];
    w.sync();
    if (leaf == T->root && T->sims_done == 0 && (T->flags & MCTS_F_NOISE)) {   // :141-143
        mcts_root_noise(w, ed, k, P, dir, P.game_base + (uint32_t)t, (uint32_t)nd->ply, dscratch);
    } else {                                                                      // normalise :144
        if (w.lane == 0) {
            float s = 0.f;
            for (int i = 0; i < k; i++) s = MC_FADD(s, ed[i].P);
            reinterpret_cast<float*>(dscratch)[0] = s;
        }
        w.sync();
        const float s = reinterpret_cast<float*>(dscratch)[0];
        for (int i = w.lane; i < k; i += W::W) ed[i].P = MC_FDIV(ed[i].P, s);
        w.sync();
    }
    float v[N];
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = vin[i];
    if (w.lane == 0) {
        nd->u.x.Ns = 0;
        nd->u.x.Qs = v[0];   // :147
        nd->kind = MCTS_NODE_EXPANDED;
    }
    w.sync();
    mcts_backup<N>(w, A, t, T->path_len, v);
    w.sync();
    if (w.lane == 0) {
        T->sims_done += 1;
        T->leaf = -1; T->cur = -1; T->pend_edge = -1; T->path_len = 0;
        T->nn_calls += 1;
    }
    w.sync();
}

// ------------------------------------------------------------------------------------------
// start of a move: find or create the root; make room first if the pools could run out
// ------------------------------------------------------------------------------------------
// re-bases the simulation in flight after the nodes moved: remap[old] = new index + 1, nodes[] already at their new places;
// path edge entries were made node-relative beforehand (mcts_path_make_relative)
template <class W>
SPL_D int mcts_path_make_relative(const W& w, const MctsArena& A, int t, int* pend_rel) {
    MctsTree* T = A.trees + t;
    const MctsNode* nodes = A.nodes + (size_t)t * A.cap;
    uint32_t* path = A.path + (size_t)t * A.max_depth * 2;
    const bool in_flight = T->leaf >= 0 || T->cur >= 0 || T->pend_edge >= 0;
    const int plen = in_flight ? T->path_len : 0;
    for (int d = w.lane; d < plen; d += W::W) path[2 * d + 1] -= nodes[path[2 * d]].edge_off;
    *pend_rel = T->pend_edge >= 0 ? T->pend_edge - (int)nodes[T->pend_parent].edge_off : -1;
    w.sync();
    return plen;
}
template <class W>
SPL_D void mcts_rebase_in_flight(const W& w, const MctsArena& A, int t, const uint32_t* remap, int plen, int pend_rel) {
    MctsTree* T = A.trees + t;
    const MctsNode* nodes = A.nodes + (size_t)t * A.cap;
    uint32_t* path = A.path + (size_t)t * A.max_depth * 2;
    for (int d = w.lane; d < plen; d += W::W) {
        const uint32_t nn = remap[path[2 * d]] - 1u;
        path[2 * d] = nn;
        path[2 * d + 1] += nodes[nn].edge_off;
    }
    if (w.lane == 0) {
        if (T->root >= 0) T->root = (int)remap[T->root] - 1;       // (-1 if the root itself was dropped: begin re-creates it)
        if (T->leaf >= 0) T->leaf = (int)remap[T->leaf] - 1;
        if (T->cur >= 0) T->cur = (int)remap[T->cur] - 1;
        if (pend_rel >= 0) {
            T->pend_parent = (int)remap[T->pend_parent] - 1;
            T->pend_edge = (int)nodes[T->pend_parent].edge_off + pend_rel;
        }
    }
    w.sync();
}

// hash table from scratch for nodes [0, n): the lanes insert concurrently (compare-and-swap on the slots)
template <class W>
SPL_D void mcts_rebuild_table(const W& w, const MctsArena& A, int t, int n) {
    uint32_t* tab = A.htab + (size_t)t * A.hcap;
    const MctsNode* nodes = A.nodes + (size_t)t * A.cap;
    for (int i = w.lane; i < A.hcap; i += W::W) tab[i] = 0u;
    w.sync();
    for (int i = w.lane; i < n; i += W::W) {
        uint32_t slot = (uint32_t)nodes[i].hash & (uint32_t)(A.hcap - 1);
        for (;;) {
#ifdef __CUDACC__
            if (atomicCAS(&tab[slot], 0u, (uint32_t)i + 1u) == 0u) break;
#else
            if (tab[slot] == 0u) { tab[slot] = (uint32_t)i + 1u; break; }
#endif
            slot = (slot + 1u) & (uint32_t)(A.hcap - 1);
        }
    }
    w.sync();
}

// In-place compaction keeping nodes with ply >= min_ply. A node below the root's ply can never be looked up again
// (every key carries its ply), so this is result-neutral - the reference's own cleaning (:80-85) relies on the same
// fact. Node indices and edge offsets both grow in creation order, so every block only moves towards the front.
template <class W>
SPL_D void mcts_compact(const W& w, const MctsArena& A, int t, int min_ply, bool use_marks) {
    MctsTree* T = A.trees + t;
    MctsNode* nodes = A.nodes + (size_t)t * A.cap;
    MctsEdge* edges = A.edges + (size_t)t * A.ecap;
    uint32_t* remap = A.htab + (size_t)t * A.hcap;   // the table is rebuilt below; hcap >= 2 cap
    int8_t* states = A.states + (size_t)t * A.cap * A.sp;
    const int n_old = T->n_nodes;
    int n_new = 0;
    for (int base = 0; base < n_old; base += W::W) {
        const int i = base + w.lane;
        const bool live = i < n_old && (use_marks ? remap[i] != 0u : (int)nodes[i].ply >= min_ply);
        const uint32_t b = w.ballot(live);
        if (i < n_old) remap[i] = live ? (uint32_t)(n_new + SPL_POPC(b & w.lanemask_lt())) + 1u : 0u;
        n_new += SPL_POPC(b);
    }
    w.sync();
    // a simulation may be in flight (the periodic cleaning runs between waves): its references into the pools - root,
    // cur, leaf, the pending edge and the recorded path - are re-based with the nodes. Edge references become
    // (node, position inside the node) while the blocks move.
    int pend_rel;
    const int plen = mcts_path_make_relative(w, A, t, &pend_rel);
    int e_new = 0;
    for (int i = 0; i < n_old; i++) {
        const uint32_t r = remap[i];
        if (r == 0u) continue;
        const int j = (int)r - 1;
        MctsNode nd = nodes[i];
        const int ne = nd.kind == MCTS_NODE_TERMINAL ? 0 : (int)nd.n_edges;
        const int off = (int)nd.edge_off;
        for (int c = 0; c < ne; c += W::W) {   // chunk-wise: read, sync, write (destination never passes the source)
            const int k = c + w.lane;
            MctsEdge e;
            if (k < ne) {
                e = edges[off + k];
                if (e.child) e.child = remap[e.child - 1u];
            }
            w.sync();
            if (k < ne) edges[e_new + k] = e;
            w.sync();
        }
        if (j != i) {
            for (int c = 0; c < A.sp / 16; c += W::W) {
                const int k = c + w.lane;
                uint4 x;
                if (k < A.sp / 16) x = reinterpret_cast<const uint4*>(states + (size_t)i * A.sp)[k];
                w.sync();
                if (k < A.sp / 16) reinterpret_cast<uint4*>(states + (size_t)j * A.sp)[k] = x;
            }
        }
        w.sync();
        if (w.lane == 0) {
            if (nd.kind != MCTS_NODE_TERMINAL) nd.edge_off = (uint32_t)e_new;
            nodes[j] = nd;
        }
        e_new += ne;
        w.sync();
    }
    for (int k = w.lane; k < e_new; k += W::W) {   // the children moved too: refresh the edge ranges cached in the edges
        MctsEdge* e = edges + k;
        if (e->child) e->child_eoff = nodes[e->child - 1u].edge_off;
    }
    w.sync();
    mcts_rebase_in_flight(w, A, t, remap, plen, pend_rel);
    mcts_rebuild_table(w, A, t, n_new);
    if (w.lane == 0) {
        T->n_nodes = n_new;
        T->n_edges = e_new;
        T->compactions += 1;
    }
    w.sync();
}

// numbers (in the hash-table region, which the compaction rebuilds anyway) every node reachable from `root` through
// linked edges in breadth-first order: mark[old index] = position + 1, queue[position] = old index; returns the count.
// Frontier nodes are expanded 32 at a time, one lane per node, edge position by edge position, so the numbering is
// deterministic. Tighter than the ply rule, but it also drops nodes that a not-yet-linked edge could still transpose into.
template <class W>
SPL_D int mcts_mark_reachable(const W& w, const MctsArena& A, int t, int root) {
    MctsTree* T = A.trees + t;
    const MctsNode* nodes = A.nodes + (size_t)t * A.cap;
    const MctsEdge* edges = A.edges + (size_t)t * A.ecap;
    uint32_t* mark = A.htab + (size_t)t * A.hcap;
    uint32_t* queue = mark + A.cap;   // hcap >= 2 cap
    const int n_old = T->n_nodes;
    for (int i = w.lane; i < n_old; i += W::W) mark[i] = i == root ? 1u : 0u;
    if (w.lane == 0) queue[0] = (uint32_t)root;
    w.sync();
    int head = 0, tail = 1;
    while (head < tail) {
        const int q = head + w.lane;
        const int level_end = tail;          // nodes numbered so far; this pass expands queue[head .. min(head + W, tail))
        int ne = 0;
        uint32_t eoff = 0u;
        if (q < level_end) {
            const MctsNode nd = nodes[queue[q]];
            ne = nd.kind == MCTS_NODE_TERMINAL ? 0 : (int)nd.n_edges;
            eoff = nd.edge_off;
        }
        int max_ne = ne;
#ifdef __CUDACC__
        max_ne = __reduce_max_sync(0xffffffffu, max_ne);
#endif
        for (int k = 0; k < max_ne; k++) {
            bool fresh = false;
            uint32_t child = 0u;
            if (k < ne) {
                child = edges[eoff + k].child;
                if (child) {
#ifdef __CUDACC__
                    fresh = atomicCAS(&mark[child - 1u], 0u, 0xFFFFFFFFu) == 0u;
#else
                    fresh = mark[child - 1u] == 0u;
                    if (fresh) mark[child - 1u] = 0xFFFFFFFFu;
#endif
                }
            }
            const uint32_t b = w.ballot(fresh);
            if (fresh) {
                const int pos = tail + SPL_POPC(b & w.lanemask_lt());
                queue[pos] = child - 1u;
                mark[child - 1u] = (uint32_t)pos + 1u;
            }
            tail += SPL_POPC(b);
            w.sync();
        }
        head = level_end < head + W::W ? level_end : head + W::W;
    }
    w.sync();
    return tail;
}

// Reachable cleaning, fast path: the reachable nodes (numbered breadth-first by mcts_mark_reachable, n_live of them) are
// copied through the FREE TAIL of the pools - out to [n_old, n_old + n_live), then back to the front - so every copy is
// independent: one lane per node, no ordering constraints (the in-place mcts_compact walks the nodes one by one).
// New node index = breadth-first position; edge blocks follow in that order, so offsets still grow with the index.
// Returns false (nothing changed) when the tails are too small; the caller then compacts in place.
template <class W>
SPL_D bool mcts_compact_reachable(const W& w, const MctsArena& A, int t, int n_live) {
    MctsTree* T = A.trees + t;
    MctsNode* nodes = A.nodes + (size_t)t * A.cap;
    MctsEdge* edges = A.edges + (size_t)t * A.ecap;
    uint32_t* mark = A.htab + (size_t)t * A.hcap;
    const uint32_t* queue = mark + A.cap;
    int8_t* states = A.states + (size_t)t * A.cap * A.sp;
    const int n_old = T->n_nodes, e_old = T->n_edges;
    if (n_old + n_live > A.cap) return false;
    int live_edges = 0;
    for (int q = w.lane; q < n_live; q += W::W) {
        const MctsNode nd = nodes[queue[q]];
        live_edges += nd.kind == MCTS_NODE_TERMINAL ? 0 : (int)nd.n_edges;
    }
    live_edges = w.sum(live_edges);
    if (e_old + live_edges > A.ecap) return false;
    int pend_rel;
    const int plen = mcts_path_make_relative(w, A, t, &pend_rel);
    // pass 1: out to the tails (headers with their new edge offsets, states, edges with re-numbered children)
    int ebase = 0;
    for (int base = 0; base < n_live; base += W::W) {
        const int q = base + w.lane;
        MctsNode nd;
        int ne = 0, old = 0;
        if (q < n_live) {
            old = (int)queue[q];
            nd = nodes[old];
            ne = nd.kind == MCTS_NODE_TERMINAL ? 0 : (int)nd.n_edges;
        }
        int incl = ne;   // inclusive scan of the edge counts over the lanes
#ifdef __CUDACC__
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (w.lane >= o) incl += v;
        }
#endif
        const int new_off = ebase + incl - ne;
        if (q < n_live) {
            const uint32_t old_off = nd.edge_off;
            if (nd.kind != MCTS_NODE_TERMINAL) nd.edge_off = (uint32_t)new_off;
            nodes[n_old + q] = nd;
            const uint4* s4 = reinterpret_cast<const uint4*>(states + (size_t)old * A.sp);
            uint4* d4 = reinterpret_cast<uint4*>(states + (size_t)(n_old + q) * A.sp);
            for (int k = 0; k < A.sp / 16; k++) d4[k] = s4[k];
            for (int k = 0; k < ne; k++) {
                MctsEdge e = edges[old_off + k];
                if (e.child) e.child = mark[e.child - 1u];
                edges[e_old + new_off + k] = e;
            }
        }
#ifdef __CUDACC__
        ebase += __shfl_sync(0xffffffffu, incl, 31);
#else
        ebase += incl;
#endif
        w.sync();
    }
    // pass 2: back to the front
    for (int q = w.lane; q < n_live; q += W::W) {
        nodes[q] = nodes[n_old + q];
        const uint4* s4 = reinterpret_cast<const uint4*>(states + (size_t)(n_old + q) * A.sp);
        uint4* d4 = reinterpret_cast<uint4*>(states + (size_t)q * A.sp);
        for (int k = 0; k < A.sp / 16; k++) d4[k] = s4[k];
    }
    w.sync();
    for (int k = w.lane; k < live_edges; k += W::W) {
        MctsEdge e = edges[e_old + k];
        if (e.child) e.child_eoff = nodes[e.child - 1u].edge_off;
        edges[k] = e;
    }
    w.sync();
    mcts_rebase_in_flight(w, A, t, mark, plen, pend_rel);
    mcts_rebuild_table(w, A, t, n_live);
    if (w.lane == 0) { T->n_nodes = n_live; T->n_edges = live_edges; T->compactions += 1; }
    w.sync();
    return true;
}

// reachable cleaning of tree t from `root`: fast path if the pools have room for it, in place otherwise
template <class W>
SPL_D void mcts_clean_reachable(const W& w, const MctsArena& A, int t, int root, bool in_place_only) {
    const int n_live = mcts_mark_reachable(w, A, t, root);
    if (in_place_only || !mcts_compact_reachable(w, A, t, n_live)) mcts_compact(w, A, t, 0, true);
}

// periodic cleaning between waves (any state of the search): trees whose pools are filled beyond the thresholds drop
// what can no longer be used - nodes below the root's ply (exact) or everything the root does not reach (gc_reachable).
// Every tree that needs it cleans in the SAME launch, so the serial per-tree work is paid once for all of them.
template <class W>
SPL_D void mcts_clean_tree(const W& w, const MctsArena& A, int t, int max_nodes, int max_edges, int gc_reachable) {
    MctsTree* T = A.trees + t;
    if (T->root < 0 || (T->n_nodes <= max_nodes && T->n_edges <= max_edges)) return;
    if (gc_reachable) mcts_clean_reachable(w, A, t, T->root, gc_reachable == 2);
    else mcts_compact(w, A, t, (int)A.nodes[(size_t)t * A.cap + T->root].ply, false);
}

template <class W>
SPL_D void mcts_clear_tree(const W& w, const MctsArena& A, int t) {   // reset_all_search_trees :188-192 for one tree
    uint32_t* tab = A.htab + (size_t)t * A.hcap;
    for (int i = w.lane; i < A.hcap; i += W::W) tab[i] = 0u;
    if (w.lane == 0) {
        MctsTree* T = A.trees + t;
        T->n_nodes = 0; T->n_edges = 0; T->root = -1; T->leaf = -1; T->sims_done = 0; T->sims_target = 0; T->path_len = 0;
        T->flags = 0u; T->status = 0u; T->depth_sum = 0; T->spec_hits = 0; T->cur = -1; T->pend_edge = -1; T->pend_parent = -1;
    }
    w.sync();
}

// root_state: the reference's int8[R,7] bytes. edge_reserve: edges budgeted per new node when deciding to clean.
// gc_reachable = 0: cleaning keeps every node with ply >= the root's (exactly result-neutral, the parity mode);
// gc_reachable = 1: keeps only what is reachable from the new root (production: far smaller pools; a node that only a
// not-yet-linked edge transposes into is re-created instead of found, which the reference's dictionary would not do).
template <int N, class W>
SPL_D void mcts_begin_tree(const W& w, const MctsArena& A, int t, const MctsSearchParams& P, const int8_t* root_state, int sims_target,
                           uint32_t flags, int edge_reserve, int gc_reachable, const double* dir, int8_t* st, uint32_t* scratch, double* dscratch) {
    typedef MctsLay<N> ML;
    MctsTree* T = A.trees + t;
    for (int i = w.lane; i < ML::SP; i += W::W) st[i] = i < ML::S ? root_state[i] : (int8_t)0;
    w.sync();
    const int root_ply = (int)(uint8_t)st[6];
    const uint64_t h = mcts_hash(w, st, A.sp);
    // room for this move: one node per simulation and, per node, the larger of the configured reserve and 1.5 x this
    // tree's own average number of legal moves
    const int avg = T->n_nodes > 16 ? (T->n_edges + T->n_nodes - 1) / T->n_nodes : 0;
    const int per_node = edge_reserve > avg + avg / 2 ? edge_reserve : avg + avg / 2;
    const int need_nodes = sims_target + 2, need_edges = (sims_target + 2) * per_node;
    const uint32_t overflowed = T->status & (MCTS_S_OVERFLOW_NODES | MCTS_S_OVERFLOW_EDGES);
    if (overflowed || T->n_nodes + need_nodes > A.cap || T->n_edges + need_edges > A.ecap) {
        bool cleared = false;
        if (gc_reachable) {
            const int old_root = mcts_lookup(w, A, t, st, h);
            if (old_root >= 0) {
                mcts_clean_reachable(w, A, t, old_root, gc_reachable == 2);
            } else {   // a root the tree has never seen (a card was revealed): nothing of the old tree can be reached
                const int resets = T->resets;
                mcts_clear_tree(w, A, t);
                if (w.lane == 0) T->resets = resets;
                w.sync();
                cleared = true;
            }
        } else {
            mcts_compact(w, A, t, root_ply, false);
            if (T->n_nodes + need_nodes > A.cap || T->n_edges + need_edges > A.ecap) {
                // the exact cleaning did not free enough: fall back to the reachable set of the new root (counted as lossy:
                // nodes that only a not-yet-linked edge could transpose into are dropped)
                const int old_root = mcts_lookup(w, A, t, st, h);
                if (old_root >= 0) {
                    mcts_clean_reachable(w, A, t, old_root, false);
                    if (w.lane == 0) T->resets += 1;
                    w.sync();
                }
            }
        }
        if (!cleared && (T->n_nodes + need_nodes > A.cap || T->n_edges + need_edges > A.ecap)) {   // still no room: forget the tree (counted)
            const int resets = T->resets;
            mcts_clear_tree(w, A, t);
            if (w.lane == 0) T->resets = resets + 1;
            w.sync();
        }
        if (w.lane == 0) {
            if (overflowed) T->truncated += 1;
            T->status &= ~(uint32_t)(MCTS_S_OVERFLOW_NODES | MCTS_S_OVERFLOW_EDGES);
        }
        w.sync();
    }
    int idx = mcts_lookup(w, A, t, st, h);
    if (idx < 0) idx = mcts_create_node<N>(w, A, t, P, st, h, scratch);
    if (w.lane == 0) {
        T->root = idx; T->leaf = -1; T->sims_done = 0; T->sims_target = idx < 0 ? 0 : sims_target; T->path_len = 0;
        T->flags = flags; T->cur = -1; T->pend_edge = -1; T->pend_parent = -1;
    }
    w.sync();
    // a root the tree already expanded gets the noise on its stored Ps before the first simulation picks (:150-154);
    // a new root gets it when its network row arrives (mcts_expand_tree)
    if (idx >= 0 && sims_target > 0 && (flags & MCTS_F_NOISE)) {
        MctsNode* nd = A.nodes + (size_t)t * A.cap + idx;
        if (nd->kind == MCTS_NODE_EXPANDED)
            mcts_root_noise(w, A.edges + (size_t)t * A.ecap + nd->edge_off, (int)nd->n_edges, P, dir, P.game_base + (uint32_t)t, (uint32_t)nd->ply, dscratch);
    }
}

// ------------------------------------------------------------------------------------------
// getActionProb's tail (:61-97) for one tree: probs double[406], q double[N]
// temp == 0 -> one-hot of the FIRST best action (the reference picks a random one among ties)
// ------------------------------------------------------------------------------------------
template <int N, class W>
SPL_D void mcts_policy_tree(const W& w, const MctsArena& A, int t, double temp, double* probs, double* q, double* dscratch) {
    const MctsTree* T = A.trees + t;
    for (int a = w.lane; a < SPL_ACTIONS; a += W::W) probs[a] = 0.0;
    w.sync();
    if (T->root < 0) return;
    const MctsNode* nd = A.nodes + (size_t)t * A.cap + T->root;
    if (nd->kind != MCTS_NODE_EXPANDED) return;
    const MctsEdge* ed = A.edges + (size_t)t * A.ecap + nd->edge_off;
    const int k = nd->n_edges;
    const bool forced = (T->flags & MCTS_F_FORCED) != 0u;
    if (w.lane == 0) {
        const float qs = nd->u.x.Qs;
        for (int p = 0; p < N; p++) q[p] = p == 0 ? (double)qs : (double)MC_FDIV(-qs, (float)(N - 1));   // :65-66
        int best = 0;
        for (int i = 0; i < k; i++) best = ed[i].N > best ? ed[i].N : best;
        double sum = 0.0;
        int besti = -1;
        double bestc = -1.0;
        for (int i = 0; i < k; i++) {
            double c = (double)ed[i].N;
            if (forced) {   // policy target pruning :69-74
                if (ed[i].N != best) c = c - (double)(long long)MC_DSQRT(MC_DMUL(MC_DMUL(MCTS_KFORCED, (double)ed[i].P), (double)T->sims_target));
                c = c > 1.0 ? c : 0.0;
            }
            if (c > bestc) { bestc = c; besti = i; }
            if (temp != 0.0) {
                c = temp == 1.0 ? c : pow(c, 1.0 / temp);   // :94
                sum = MC_DADD(sum, c);
            }
            probs[ed[i].action] = c;
        }
        if (temp == 0.0) {   // :87-92
            for (int i = 0; i < k; i++) probs[ed[i].action] = i == besti ? 1.0 : 0.0;
        } else {
            for (int i = 0; i < k; i++) probs[ed[i].action] = MC_DDIV(probs[ed[i].action], sum);   // :95-96
        }
    }
    (void)dscratch;
    w.sync();
}

// getActionProb's tail + the caller's pick from it, for a tree whose budget is spent (Coach.py:75-86: `pi = getActionProb(...)`,
// `action = np.random.choice(len(pi), p=pi)`), without materialising the 406 probabilities: the weights of mcts_policy_tree
// (counts -> policy-target pruning -> temperature), then one uniform from Philox keyed (seed, game, episode, root ply) and a walk
// through the cumulative weights in action order. Returns the action, or -1 while the tree's search is still running.
// temp == 0: the first most visited action. A finished tree without a single visit (never in a game in progress) gives -1 too.
template <int N, class W>
SPL_D int mcts_sample_tree(const W& w, const MctsArena& A, int t, const MctsSearchParams& P, double temp, uint32_t episode, bool* finished) {
    const MctsTree* T = A.trees + t;
    const bool fin = T->root >= 0 && (T->sims_done >= T->sims_target || T->status != 0u);
    *finished = fin;
    if (!fin) return -1;
    const MctsNode* nd = A.nodes + (size_t)t * A.cap + T->root;
    if (nd->kind != MCTS_NODE_EXPANDED) return -1;
    const MctsEdge* ed = A.edges + (size_t)t * A.ecap + nd->edge_off;
    const int k = nd->n_edges;
    const bool forced = (T->flags & MCTS_F_FORCED) != 0u;
    int action = -1;
    if (w.lane == 0) {
        int best = 0;
        for (int i = 0; i < k; i++) best = ed[i].N > best ? ed[i].N : best;
        double sum = 0.0, bestc = -1.0;
        int besti = -1;
        for (int pass = 0; pass < 2; pass++) {      // pass 0: total weight, pass 1: the walk up to u * total
            double target = 0.0, acc = 0.0;
            if (pass == 1) {
                if (temp == 0.0 || !(sum > 0.0)) break;
                const SplPhilox r = spl_philox(P.seed, P.game_base + (uint32_t)t, episode, (uint32_t)nd->ply, 5);
                target = MC_DMUL(mcts_u01(r.v[0], r.v[1]), sum);
            }
            for (int i = 0; i < k; i++) {
                double c = (double)ed[i].N;
                if (forced) {   // policy target pruning :69-74
                    if (ed[i].N != best) c = c - (double)(long long)MC_DSQRT(MC_DMUL(MC_DMUL(MCTS_KFORCED, (double)ed[i].P), (double)T->sims_target));
                    c = c > 1.0 ? c : 0.0;
                }
                if (pass == 0 && c > bestc) { bestc = c; besti = i; }
                if (temp != 0.0 && temp != 1.0) c = pow(c, 1.0 / temp);   // :94
                if (pass == 0) sum = MC_DADD(sum, c);
                else {
                    acc = MC_DADD(acc, c);
                    if (c > 0.0) action = (int)ed[i].action;      // the last action with weight, should rounding leave acc <= target at the end
                    if (acc > target) break;
                }
            }
        }
        if (temp == 0.0 && besti >= 0 && bestc > 0.0) action = (int)ed[besti].action;
    }
    action = w.shfl(action, 0);
    return action;
}

// raw root statistics (tests, diagnostics): Nsa int32[406], Qsa double[406], Ps float[406], info int32[16]
template <class W>
SPL_D void mcts_root_stats_tree(const W& w, const MctsArena& A, int t, int32_t* nsa, double* qsa, float* ps, int32_t* info) {
    const MctsTree* T = A.trees + t;
    for (int a = w.lane; a < SPL_ACTIONS; a += W::W) {
        if (nsa) nsa[a] = 0;
        if (qsa) qsa[a] = MCTS_UNVISITED;
        if (ps) ps[a] = 0.f;
    }
    w.sync();
    int ns = 0;
    float qs = 0.f;
    if (T->root >= 0) {
        const MctsNode* nd = A.nodes + (size_t)t * A.cap + T->root;
        if (nd->kind == MCTS_NODE_EXPANDED) {
            const MctsEdge* ed = A.edges + (size_t)t * A.ecap + nd->edge_off;
            for (int i = w.lane; i < (int)nd->n_edges; i += W::W) {
                if (nsa) nsa[ed[i].action] = ed[i].N;
                if (qsa) qsa[ed[i].action] = ed[i].Q;
                if (ps) ps[ed[i].action] = ed[i].P;
            }
            ns = nd->u.x.Ns; qs = nd->u.x.Qs;
        }
    }
    if (info && w.lane == 0) {
        info[0] = T->n_nodes; info[1] = T->n_edges; info[2] = ns; info[3] = T->sims_done; info[4] = T->nn_calls;
        info[5] = (int32_t)(T->status | ((uint32_t)T->truncated << 8)); info[6] = T->resets * 65536 + T->compactions;
        memcpy(&info[7], &qs, 4);
        memcpy(&info[8], T->last_v, 16);
        info[12] = T->depth_sum;
        info[13] = T->spec_hits;
    }
    w.sync();
}

// ------------------------------------------------------------------------------------------
// deterministic stand-in network ("fixed NN outputs"): policy and values are a pure function of the state bytes,
// all probabilities multiples of 2^-13 that sum to exactly 1 (tests/golden/mcts_*.npz were produced by the
// reference's own MCTS.py with the same function as its network)
// ------------------------------------------------------------------------------------------
template <int N, class W>
SPL_D void mcts_fixed_net_row(const W& w, const int8_t* state, const uint8_t* valid, float* pi, float* v, uint32_t* scratch) {
    typedef MctsLay<N> ML;
    if (w.lane == 0) {
        uint64_t h = 0xCBF29CE484222325ull;   // FNV-1a 64 over the state bytes
        for (int i = 0; i < ML::S; i++) { h ^= (uint64_t)(uint8_t)state[i]; h *= 0x100000001B3ull; }
        scratch[0] = (uint32_t)h; scratch[1] = (uint32_t)(h >> 32);
    }
    w.sync();
    const uint64_t h = (uint64_t)scratch[0] | ((uint64_t)scratch[1] << 32);
    int part = 0, first = SPL_ACTIONS;
    for (int a = w.lane; a < SPL_ACTIONS; a += W::W) {
        float p = 0.f;
        if (valid[a]) {
            const int wgt = 1 + (int)(mcts_mix64(h + (uint64_t)a * 0x9E3779B97F4A7C15ull) >> 58);
            p = (float)wgt;
            part += wgt;
            if (first == SPL_ACTIONS) first = a;
        }
        pi[a] = p;
    }
    const int total = w.sum(part);
#ifdef __CUDACC__
    first = __reduce_min_sync(0xffffffffu, first);
#endif
    w.sync();
    if (w.lane == 0 && first < SPL_ACTIONS) pi[first] += (float)(8192 - total);
    w.sync();
    for (int a = w.lane; a < SPL_ACTIONS; a += W::W) pi[a] = pi[a] / 8192.f;
    for (int p = w.lane; p < N; p += W::W) v[p] = (float)((double)((long long)((h >> (8 * p + 3)) & 0x7Full) - 64) / 64.0);
    w.sync();
}
