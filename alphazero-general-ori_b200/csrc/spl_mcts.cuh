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
// transpositions that persists across moves) becomes a per-tree open-addressing hash table (64-bit state hash + full-state
// compare) over variable-size node RECORDS that all trees allocate from ONE shared pool of 32 KB pages: a tree holds what
// it needs at the moment (a few hundred nodes after a card was revealed, tens of thousands in a long line without reveals)
// instead of a worst-case slab. A record = header + the reference's state bytes + the node's sparse edges (legal actions
// only, in action order, structure-of-arrays: Q[k] | {P, N}[k] | {child, action, child's edge count}[k]), so one pointer
// names a node and its header and edges arrive in one round trip. Nodes are TERMINAL (Es stored), NEEDS_NN (edges
// allocated, waiting for the network) or EXPANDED. An inner traversal touches no rules code at all.
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
    __device__ __forceinline__ int scan_excl(int v, int& total) const {   // exclusive prefix sum over the lanes
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        total = __shfl_sync(0xffffffffu, incl, 31);
        return incl - v;
    }
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
    int scan_excl(int v, int& total) const { total = v; return 0; }
    void best(double&, int&) const {}
    void sqrt2(double a, double b, double& ra, double& rb) const { ra = MC_DSQRT(a); rb = MC_DSQRT(b); }
};
#endif


// ------------------------------------------------------------------------------------------
// arena layout (all in caller-owned HBM)
// ------------------------------------------------------------------------------------------
enum { MCTS_NODE_TERMINAL = 1, MCTS_NODE_NEEDS_NN = 2, MCTS_NODE_EXPANDED = 3 };
enum { MCTS_F_FORCED = 1u, MCTS_F_NOISE = 2u };                       // per-move flags (getActionProb :54-58)
enum { MCTS_S_OVERFLOW_NODES = 1u, MCTS_S_OVERFLOW_POOL = 2u, MCTS_S_PROTOCOL = 4u, MCTS_S_BAD_STATE = 8u };   // sticky status bits

#define MCTS_UNIT 32u            // the pool is addressed in 32-byte units: a 32-bit record offset spans 128 GB
#define MCTS_PAGE_UNITS 1024u    // 32 KB pages; the largest record (4 players, 406 edges: 10.4 KB) fits

struct MctsNode {   // 32 B record header
    uint64_t hash;
    uint8_t ply, kind;
    uint16_t n_edges;
    uint32_t fwd;   // 0 outside a compaction; during one: the record's new offset (old copy) / the next record of the lane's chain (new copy)
    union {
        struct { int32_t Ns; float Qs; uint32_t hint; uint32_t pad; } x;   // EXPANDED / NEEDS_NN; hint: see mcts_descend_tree
        float es[4];                                                       // TERMINAL: getGameEnded vector
    } u;
};
struct MctsPN { float P; int32_t N; };                                     // Ps[a], Nsa
struct MctsCA { uint32_t child; uint16_t action; uint16_t child_ne; };     // child record (0 = not linked yet), action, the child's edge count
struct MctsEdgeV {   // one edge in registers
    double Q;        // Qsa (float64, -42 = unvisited)
    float P;
    int32_t N;
    uint32_t child;
    uint16_t action, child_ne;
};
#define MCTS_KMAX 4   // simulations one tree may have in flight (leaves_per_tree; 1 = the reference's sequential search)
struct MctsSlot {     // state of one simulation in flight (it survives between the kernels of a wave)
    uint32_t leaf;         // != 0: this node waits for the network
    uint32_t cur;          // != 0: continue the descent at this node with path_len edges already recorded; 0: start at the root
    uint32_t pend_parent;
    int32_t pend_edge;     // >= 0: the descent stopped at this edge of pend_parent (position | edge count << 16), whose child state is being computed / attached
    int32_t path_len;
};
struct MctsTree {   // 192 B
    int32_t n_nodes, n_edges;          // records / edges this tree holds
    uint32_t root;                     // record offset, 0 = none
    int32_t sims_done, sims_target;
    uint32_t flags, status;
    int32_t nn_calls, resets, compactions;
    float last_v[4];   // value vector the last finished simulation returned at the root (what MCTS.search returns)
    int32_t truncated; // searches cut short because the tree hit its node limit or the pool ran dry mid-move
    int32_t depth_sum; // sum of path lengths of the simulations since the last reset (diagnostics)
    int32_t spec_hits;    // diagnostics: levels of the descents that started from the early-fetched child (see mcts_descend_tree)
    MctsSlot slot[MCTS_KMAX];
    // storage
    int32_t n_pages;      // pages owned: tree_pages[t][0 .. n_pages)
    uint32_t bump, page_end;   // next free unit / end of the page being filled
    int32_t n_dropped;    // nodes an exact cleaning dropped since the last reset: n_nodes + n_dropped = size of the reference's dictionary
    uint32_t episode;     // the lane's episode counter at the last begin (Philox key of the on-device Dirichlet sampler)
    uint8_t deck[15];     // the deck bitmasks (rows 26 / 28 / 30, five colours each) every node of a homogeneous tree carries
    uint8_t hetero;       // 1: nodes with different decks may be present (roots that do not follow each other in one game)
    int32_t pad[2];
};
struct MctsArena {
    int n_trees, node_limit, hcap, sp, cp, max_depth, max_pages;   // sp: padded bytes of the reference's state (staging rows); cp: of its compact form (records)
    int n_slots;            // leaves_per_tree: 1 = parity mode; > 1: virtual-loss leaf batching (rows of the per-tree buffers below: tree * n_slots + slot)
    uint32_t n_pool_pages;
    uint8_t* pool;          // [n_pool_pages][32 KB]   node records of all trees (page 0 is never handed out: offset 0 = none)
    uint32_t* fq_slots;     // [n_pool_pages]          ring of free pages (0 = empty slot)
    int32_t* fq_ctl;        // [0] head ticket, [1] tail ticket, [2] free pages, [3] fewest free pages seen (diagnostics)
    uint32_t* tree_pages;   // [T][max_pages]          pages of every tree; the upper half is room for the copy of a compaction
    uint32_t* htab;         // [T][hcap]               record offsets, linear probing
    MctsTree* trees;        // [T]
    uint32_t* path;         // [T][n_slots][max_depth][2]  (record, edge position | edge count << 16) of the simulations in flight
    // staging between the rules kernel and the attach kernel (the child state of every tree's pending edge)
    int8_t* stage_state;   // [T * n_slots][sp]
    uint32_t* stage_mask;  // [13][T * n_slots]
    float* stage_es;       // [T * n_slots][4]
    uint8_t* stage_ended;  // [T * n_slots]
    long long* prof;       // diagnostics (normally NULL): [T][16] per-tree time stamps of the last wave (spl_mcts_debug_profile)
    uint8_t* leaf_src;     // [T * n_slots]  where the network input of the tree's leaf lives: 0 = its leaf row (bytes), 1 = the staging row
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

// record accessors: [MctsNode 32][compact state cp][Q double[k]][MctsPN[k]][MctsCA[k]], padded to a multiple of 32 bytes
SPL_D uint32_t* mcts_path(const MctsArena& A, int t, int s) {   // (rows x 2 max_depth words stay far below 2^32: 32-bit index arithmetic)
    return A.path + (uint32_t)(t * A.n_slots + s) * (uint32_t)(2 * A.max_depth);
}
SPL_D int mcts_row(const MctsArena& A, int t, int s) { return t * A.n_slots + s; }
// with several simulations in flight (virtual loss) the top byte of an edge's N counts the simulations currently below it
#define MCTS_VL_ONE 0x01000000
#define MCTS_N_MASK 0x00FFFFFF
SPL_D uint8_t* mcts_ptr(const MctsArena& A, uint32_t rec) { return A.pool + (size_t)rec * MCTS_UNIT; }
SPL_D MctsNode* mcts_node(const MctsArena& A, uint32_t rec) { return reinterpret_cast<MctsNode*>(mcts_ptr(A, rec)); }
SPL_D uint8_t* mcts_cstate(const MctsArena& A, uint32_t rec) { return mcts_ptr(A, rec) + 32; }
SPL_D uint32_t mcts_rec_units(const MctsArena& A, int k) { return (uint32_t)(32 + A.cp + 24 * k + 31) / MCTS_UNIT; }
struct MctsEdges {   // the three edge arrays of one record
    double* Q;
    MctsPN* pn;
    MctsCA* ca;
    SPL_M MctsEdgeV load(int i) const {
        MctsEdgeV e;
        const MctsPN p = pn[i];
        const MctsCA c = ca[i];
        e.Q = Q[i]; e.P = p.P; e.N = p.N; e.child = c.child; e.action = c.action; e.child_ne = c.child_ne;
        return e;
    }
};
SPL_D MctsEdges mcts_edges(const MctsArena& A, uint32_t rec, int k) {
    uint8_t* b = mcts_ptr(A, rec) + 32 + A.cp;
    MctsEdges e;
    e.Q = reinterpret_cast<double*>(b);
    e.pn = reinterpret_cast<MctsPN*>(b + 8 * (size_t)k);
    e.ca = reinterpret_cast<MctsCA*>(b + 16 * (size_t)k);
    return e;
}
SPL_D MctsEdgeV mcts_edge_none() {
    MctsEdgeV e;
    e.Q = MCTS_UNVISITED; e.P = 0.f; e.N = 0; e.child = 0u; e.action = 0; e.child_ne = 0;
    return e;
}

SPL_D uint64_t mcts_mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    return x;
}

// 64-bit hash of a padded state (sp bytes, 8-byte aligned): a sum of position-salted word mixes, so that any
// partition of the words over lanes gives the same value
template <class W>
SPL_D uint64_t mcts_hash(const W& w, const uint8_t* st, int sp) {
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

// ------------------------------------------------------------------------------------------
// compact node states. A record does not hold the reference's int8[R,7] array (392 / 497 / 616 bytes) but a lossless
// 80 / 96 / 128-byte form of it: the cells that carry information verbatim (bank + ply counter, the 15 deck bitmasks, the
// players' gems and card bonuses) and one byte per card / noble row pair (its index in the reference's tables,
// SplendorLogic.py:320-473) - SURVEY App. A. The transposition key is the compact form: two states are equal iff their compact
// forms are, because the encoder refuses (returns false) anything it could not give back byte for byte - a row that is no
// card of the tables, a deck count that is not the popcount of its mask, a cell the rules never write. The rules step and the
// network read the reference's layout again (mcts_decode).
// ------------------------------------------------------------------------------------------
template <int N> struct MctsCLay {
    static constexpr int BANK = 0, VIS = 7, DECK = 19, NOB = 34, PGEMS = NOB + N + 1, PNOB = PGEMS + 6 * N, PCARDS = PNOB + N * (N + 1),
                         RES = PCARDS + 6 * N, BYTES = RES + 3 * N;
    static constexpr int CP = (BYTES + 15) / 16 * 16;                       // 80 / 96 / 128
    static constexpr int ITEMS = 12 + 3 * N + (N + 1) + N * (N + 1);        // card / noble row groups: one byte each
};
#define MCTS_C_DECK 19   // offset of the 15 deck bitmasks (the same for every player count)

SPL_D uint32_t mcts_pack_card(const int8_t* cost, const int8_t* gain, bool* ok) {   // -> the packed form of spl_tables.cuh, 0 = empty slot
    uint32_t pk = 0u;
    int ones = 0, col = 0, sum = 0;
    bool good = cost[5] == 0 && cost[6] == 0 && gain[5] == 0;
    for (int c = 0; c < 5; c++) {
        good &= cost[c] >= 0 && cost[c] <= 15 && (gain[c] == 0 || gain[c] == 1);
        pk |= (uint32_t)(cost[c] & 15) << (4 * c);
        if (gain[c] == 1) { ones++; col = c; }
        sum += cost[c];
    }
    if (sum == 0 && ones == 0 && gain[6] == 0) { *ok = good; return 0u; }
    good &= ones == 1 && gain[6] >= 0 && gain[6] <= 15;
    *ok = good;
    return pk | ((uint32_t)col << 20) | ((uint32_t)(gain[6] & 15) << 24);
}
// index tier * 40 + deck colour * 8 + idx of the table entry equal to pk, or -1. The tables are grouped by the DECK colour of
// SplendorLogic.py (np_all_cards_*), which is a fixed permutation of the card's gain colour: only that group of 8 is searched.
SPL_D int mcts_find_card(uint32_t pk, int tier_lo, int tier_hi) {
    const uint32_t* tab = &SPL_CARDS[0][0][0];
    const int g = (int)((pk >> 20) & 7u);
    const int dc = g == 0 ? 3 : g == 1 ? 0 : g == 2 ? 4 : g == 3 ? 1 : 2;
    int f = -1;
    for (int t = tier_hi - 1; t >= tier_lo; t--) {   // branch-free: the 8 loads of a group are independent (one round trip)
        const uint32_t* grp = tab + 40 * t + 8 * dc;
#pragma unroll
        for (int i = 7; i >= 0; i--) f = grp[i] == pk ? 40 * t + 8 * dc + i : f;
    }
    return f;
}

// compact byte i (one of the verbatim ones) <-> cell of the reference's layout; returns false for the card / noble bytes and the padding
template <int N>
SPL_D bool mcts_cmap(int i, int* cell, int* count_cell) {
    typedef SplLay<N> L;
    typedef MctsCLay<N> C;
    *count_cell = -1;
    if (i < 7) { *cell = i; return true; }                                                              // bank, gold, ply counter
    if (i >= C::DECK && i < C::DECK + 15) {
        const int j = i - C::DECK, tier = j / 5, c = j - 5 * tier;
        *cell = 7 * (L::DECK + 2 * tier + 1) + c; *count_cell = 7 * (L::DECK + 2 * tier) + c;             // mask byte; its count = popcount
        return true;
    }
    if (i >= C::PGEMS && i < C::PGEMS + 6 * N) { const int j = i - C::PGEMS, p = j / 6; *cell = 7 * (L::PGEMS + p) + (j - 6 * p); return true; }
    if (i >= C::PCARDS && i < C::PCARDS + 6 * N) {
        const int j = i - C::PCARDS, p = j / 6, c = j - 6 * p;
        *cell = 7 * (L::PCARDS + p) + (c < 5 ? c : 6);                                                  // five bonuses + points
        return true;
    }
    return false;
}

// aos: the reference's int8[R,7] bytes (any memory); cst: CP bytes (zero padded). Every lane of the group takes part; the return value
// (the same in every lane) says whether the state is one the compact form can give back exactly. VALIDATE = false: for states the
// rules code itself produced from an accepted state (children inside the tree): only the table look-ups, no checks.
template <int N, bool VALIDATE, class W>
SPL_D bool mcts_encode(const W& w, const int8_t* aos, uint8_t* cst) {
    typedef SplLay<N> L;
    typedef MctsCLay<N> C;
    bool ok = true;
    for (int i = w.lane; i < C::CP; i += W::W) {
        int cell, cnt;
        if (mcts_cmap<N>(i, &cell, &cnt)) {
            const int v = aos[cell];
            cst[i] = (uint8_t)v;
            if (VALIDATE && cnt >= 0) ok &= (int)aos[cnt] == SPL_POPC((uint32_t)(uint8_t)v);
        } else if (i >= C::BYTES) cst[i] = 0;
    }
    if (VALIDATE) {   // cells the rules never write must be zero (the decoder writes zeros there)
        for (int i = w.lane; i < 6 + 2 * N; i += W::W) {
            const int row = i < 6 ? L::DECK + i : (i < 6 + N ? L::PGEMS + (i - 6) : L::PCARDS + (i - 6 - N));
            if (i < 6) ok &= aos[7 * row + 5] == 0 && aos[7 * row + 6] == 0;
            else if (i < 6 + N) ok &= aos[7 * row + 6] == 0;
            else ok &= aos[7 * row + 5] == 0;
        }
    }
    for (int it = w.lane; it < C::ITEMS; it += W::W) {
        int at, id = 0xFF;
        bool good = true;
        if (it < 12 + 3 * N) {   // a card: two rows (cost, gain)
            const bool vis = it < 12;
            const int row = vis ? L::CARDS + 2 * it : L::PRES + 2 * (it - 12);
            at = vis ? C::VIS + it : C::RES + (it - 12);
            const uint32_t pk = mcts_pack_card(aos + 7 * row, aos + 7 * (row + 1), &good);
            if (pk) {
                const int f = vis ? mcts_find_card(pk, it / 4, it / 4 + 1) : mcts_find_card(pk, 0, 3);
                good &= f >= 0;
                id = vis ? f - 40 * (it / 4) : f;
            }
        } else {                 // a noble row: on the table or in a player's block
            const bool tab = it < 12 + 3 * N + N + 1;
            const int j = it - (12 + 3 * N) - (tab ? 0 : N + 1);
            const int8_t* r = aos + 7 * ((tab ? L::NOBLES : L::PNOBLES) + j);
            at = (tab ? C::NOB : C::PNOB) + j;
            uint32_t pk = 0u;
            int sum = 0;
            good = r[5] == 0;
            for (int c = 0; c < 5; c++) { good &= r[c] >= 0 && r[c] <= 15; pk |= (uint32_t)(r[c] & 15) << (4 * c); sum += r[c]; }
            if (sum == 0) good &= r[6] == 0;
            else {
                good &= r[6] == 3;
                int f = -1;
#pragma unroll
                for (int q = 9; q >= 0; q--) f = SPL_NOBLES[q] == pk ? q : f;
                good &= f >= 0;
                id = f;
            }
        }
        ok &= good;
        cst[at] = (uint8_t)id;
    }
    w.sync();
    return VALIDATE ? w.ballot(!ok) == 0u : true;
}

// cst -> the reference's int8[R,7] bytes (S of them; any memory). ZEROED: the destination is already all zero.
template <int N, class W>
SPL_D void mcts_decode(const W& w, const uint8_t* cst, int8_t* aos, bool zeroed = false) {
    typedef SplLay<N> L;
    typedef MctsCLay<N> C;
    if (!zeroed) {
        for (int i = w.lane; i < L::CELLS; i += W::W) aos[i] = 0;
        w.sync();
    }
    for (int i = w.lane; i < C::BYTES; i += W::W) {   // the verbatim cells
        int cell, cnt;
        if (mcts_cmap<N>(i, &cell, &cnt)) {
            const int v = cst[i];
            aos[cell] = (int8_t)v;
            if (cnt >= 0) aos[cnt] = (int8_t)SPL_POPC((uint32_t)v);
        }
    }
    for (int it = w.lane; it < C::ITEMS; it += W::W) {
        if (it < 12 + 3 * N) {
            const bool vis = it < 12;
            const int id = cst[vis ? C::VIS + it : C::RES + (it - 12)];
            if (id == 0xFF) continue;
            const int row = vis ? L::CARDS + 2 * it : L::PRES + 2 * (it - 12);
            const uint32_t pk = (&SPL_CARDS[0][0][0])[vis ? 40 * (it / 4) + id : id];
            int8_t* r = aos + 7 * row;
            for (int c = 0; c < 5; c++) r[c] = (int8_t)((pk >> (4 * c)) & 15u);
            r[7 + (int)((pk >> 20) & 7u)] = 1;
            r[13] = (int8_t)((pk >> 24) & 15u);
        } else {
            const bool tab = it < 12 + 3 * N + N + 1;
            const int j = it - (12 + 3 * N) - (tab ? 0 : N + 1);
            const int id = cst[(tab ? C::NOB : C::PNOB) + j];
            if (id == 0xFF) continue;
            const uint32_t pk = SPL_NOBLES[id];
            int8_t* r = aos + 7 * ((tab ? L::NOBLES : L::PNOBLES) + j);
            for (int c = 0; c < 5; c++) r[c] = (int8_t)((pk >> (4 * c)) & 15u);
            r[6] = 3;
        }
    }
    w.sync();
}

// ------------------------------------------------------------------------------------------
// the shared page pool: a ring of free pages with tickets. A pop takes a ticket and then the page in that slot (the push
// that owns the slot may still be in flight: the pop waits for it); a push fills slots first and publishes the count last.
// ------------------------------------------------------------------------------------------
SPL_D uint32_t mcts_page_pop(const MctsArena& A) {   // one thread; 0 = the pool is dry
#ifdef __CUDACC__
    const int a = atomicSub(&A.fq_ctl[2], 1);
    if (a <= 0) { atomicAdd(&A.fq_ctl[2], 1); return 0u; }
    atomicMin(&A.fq_ctl[3], a - 1);
    const uint32_t ticket = atomicAdd(reinterpret_cast<unsigned int*>(&A.fq_ctl[0]), 1u);
    unsigned int* slot = A.fq_slots + ticket % A.n_pool_pages;
    uint32_t p;
    while ((p = atomicExch(slot, 0u)) == 0u) {}
    return p;
#else
    if (A.fq_ctl[2] <= 0) return 0u;
    A.fq_ctl[2] -= 1;
    if (A.fq_ctl[2] < A.fq_ctl[3]) A.fq_ctl[3] = A.fq_ctl[2];
    const uint32_t ticket = (uint32_t)A.fq_ctl[0]++;
    uint32_t* slot = A.fq_slots + ticket % A.n_pool_pages;
    const uint32_t p = *slot;
    *slot = 0u;
    return p;
#endif
}
template <class W>
SPL_D void mcts_pages_push(const W& w, const MctsArena& A, const uint32_t* pages, int n) {   // whole warp
    if (n <= 0) return;
    uint32_t base = 0u;
#ifdef __CUDACC__
    if (w.lane == 0) base = atomicAdd(reinterpret_cast<unsigned int*>(&A.fq_ctl[1]), (unsigned int)n);
    base = (uint32_t)w.shfl((int)base, 0);
    for (int i = w.lane; i < n; i += W::W) atomicExch(A.fq_slots + (base + (uint32_t)i) % A.n_pool_pages, pages[i]);
    __threadfence();
    w.sync();
    if (w.lane == 0) atomicAdd(&A.fq_ctl[2], n);
#else
    base = (uint32_t)A.fq_ctl[1];
    A.fq_ctl[1] += n;
    for (int i = 0; i < n; i++) A.fq_slots[(base + (uint32_t)i) % A.n_pool_pages] = pages[i];
    A.fq_ctl[2] += n;
#endif
    w.sync();
}

// room for a record of `units` in tree t (lane 0 allocates, every lane gets the offset); 0 = no room (status bit set).
// Pages beyond max_pages / 2 are only handed to the copy of a compaction (`for_copy`).
template <class W>
SPL_D uint32_t mcts_alloc(const W& w, const MctsArena& A, int t, uint32_t units, bool for_copy = false) {
    MctsTree* T = A.trees + t;
    uint32_t off = 0u;
    if (w.lane == 0) {
        uint32_t bump = T->bump;
        if (bump == 0u || bump + units > T->page_end) {
            bump = 0u;
            if (T->n_pages < (for_copy ? A.max_pages : A.max_pages / 2)) {
                const uint32_t p = mcts_page_pop(A);
                if (p) {
                    A.tree_pages[(size_t)t * A.max_pages + T->n_pages] = p;
                    T->n_pages += 1;
                    bump = p * MCTS_PAGE_UNITS;
                    T->page_end = bump + MCTS_PAGE_UNITS;
                }
            }
        }
        if (bump) { off = bump; T->bump = bump + units; }
        else if (!for_copy) T->status |= MCTS_S_OVERFLOW_POOL;
    }
    off = (uint32_t)w.shfl((int)off, 0);
    return off;
}

// nodes_data.get(s) :120 -> record offset or 0
template <class W>
SPL_D uint32_t mcts_lookup(const W& w, const MctsArena& A, int t, const uint8_t* st, uint64_t h) {
    const uint32_t* tab = A.htab + (size_t)t * A.hcap;
    uint32_t slot = (uint32_t)h & (uint32_t)(A.hcap - 1);
    for (;;) {
        const uint32_t e = tab[slot];
        if (e == 0u) return 0u;
        if (mcts_node(A, e)->hash == h && mcts_equal16(w, mcts_cstate(A, e), st, A.cp)) return e;
        slot = (slot + 1u) & (uint32_t)(A.hcap - 1);
    }
}

template <class W>
SPL_D void mcts_table_insert(const W& w, const MctsArena& A, int t, uint32_t rec, uint64_t h) {
    if (w.lane == 0) {
        uint32_t* tab = A.htab + (size_t)t * A.hcap;
        uint32_t slot = (uint32_t)h & (uint32_t)(A.hcap - 1);
        while (tab[slot] != 0u) slot = (slot + 1u) & (uint32_t)(A.hcap - 1);
        tab[slot] = rec;
    }
    w.sync();
}

// New node for the state with compact form `st` (cp bytes, zero padded) whose end-of-game vector / legal mask are already known:
// stores the bytes, allocates one edge per legal action (in action order), links it into the hash table.
// m: the 13 mask words at m[i * mstride]; es: N floats (used when ended). Returns the record or 0 (no room: status bit set).
template <int N, class W>
SPL_D uint32_t mcts_store_node(const W& w, const MctsArena& A, int t, const uint8_t* st, uint64_t h, bool ended, const float* es,
                               const uint32_t* m, int mstride) {
    MctsTree* T = A.trees + t;
    if (T->n_nodes >= A.node_limit) {
        if (w.lane == 0) T->status |= MCTS_S_OVERFLOW_NODES;
        w.sync();
        return 0u;
    }
    int k = 0;
    if (!ended)
        for (int i = 0; i < SPL_MASK_WORDS; i++) k += SPL_POPC(m[i * mstride]);
    const uint32_t rec = mcts_alloc(w, A, t, mcts_rec_units(A, k));
    if (rec == 0u) { w.sync(); return 0u; }
    MctsNode* nd = mcts_node(A, rec);
    if (w.lane == 0) {
        nd->hash = h;
        nd->ply = (uint8_t)st[6];
        nd->n_edges = (uint16_t)k;
        nd->fwd = 0u;
        if (ended) {
            nd->kind = MCTS_NODE_TERMINAL;
            for (int i = 0; i < 4; i++) nd->u.es[i] = i < N ? es[i] : 0.f;
        } else {
            nd->kind = MCTS_NODE_NEEDS_NN;
            nd->u.x.Ns = 0; nd->u.x.Qs = 0.f; nd->u.x.hint = 0u; nd->u.x.pad = 0u;
        }
    }
    mcts_copy16(w, mcts_cstate(A, rec), st, A.cp);
    if (!ended) {   // edges in action order: word i of the mask owns a contiguous run
        const MctsEdges ed = mcts_edges(A, rec, k);
        for (int i = w.lane; i < SPL_MASK_WORDS; i += W::W) {
            int off = 0;
            for (int j = 0; j < i; j++) off += SPL_POPC(m[j * mstride]);
            uint32_t bits = m[i * mstride];
            while (bits) {
                const int b = SPL_FFS(bits) - 1;
                bits &= bits - 1u;
                MctsPN pn; pn.P = 0.f; pn.N = 0;
                MctsCA ca; ca.child = 0u; ca.action = (uint16_t)(32 * i + b); ca.child_ne = 0;
                ed.Q[off] = MCTS_UNVISITED; ed.pn[off] = pn; ed.ca[off] = ca;
                off++;
            }
        }
    }
    w.sync();
    if (w.lane == 0) {
        T->n_nodes += 1;
        T->n_edges += k;
    }
    mcts_table_insert(w, A, t, rec, h);
    return rec;
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
SPL_D uint32_t mcts_create_node(const W& w, const MctsArena& A, int t, const MctsSearchParams& P, int8_t* st, const uint8_t* cst, uint64_t h, uint32_t* scratch) {
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
    const uint32_t rec = mcts_store_node<N>(w, A, t, cst, h, scratch[13] != 0u, es, scratch, 1);
    w.sync();
    return rec;
}

// ------------------------------------------------------------------------------------------
// root softmax + Dirichlet noise (softmax :244-250, applyDirNoise :180-186, normalise :239-242)
// ------------------------------------------------------------------------------------------
// Marsaglia-Tsang gamma(alpha,1) for the on-device Dirichlet sampler (production; parity runs inject the vector).
// Philox key (seed), counter (game, episode, root ply, 16 + 128 * edge rank + 2 * attempt [+ 1]): fresh noise for every move of every
// episode of every lane (rng.dirichlet per move, :181); streams 0..5 of the same (game, episode, ply) belong to the env and the move sampler.
SPL_D double mcts_u01(uint32_t a, uint32_t b) { return ((double)(((uint64_t)a << 21) ^ (uint64_t)(b >> 11)) + 0.5) * (1.0 / 9007199254740992.0); }
SPL_D double mcts_gamma(double alpha, uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply, uint32_t k) {
    const double a1 = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = a1 - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double g = d;
    for (uint32_t attempt = 0; attempt < 64u; attempt++) {
        const SplPhilox r0 = spl_philox(seed, game, episode, ply, 16u + k * 128u + 2u * attempt);
        const SplPhilox r1 = spl_philox(seed, game, episode, ply, 16u + k * 128u + 2u * attempt + 1u);
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
SPL_D void mcts_root_noise(const W& w, MctsPN* ed, int k, const MctsSearchParams& P, const double* dir, uint32_t game, uint32_t episode, uint32_t ply,
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
        for (int i = w.lane; i < k; i += W::W) part = MC_DADD(part, mcts_gamma(P.dirichlet_alpha, P.seed, game, episode, ply, (uint32_t)i));
        gsum = w.sumd_tree(part);
    }
    for (int i = w.lane; i < k; i += W::W) {
        const double dv = dir ? dir[i] : MC_DDIV(mcts_gamma(P.dirichlet_alpha, P.seed, game, episode, ply, (uint32_t)i), gsum);
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
// returns the edge position inside the node.
// VL (leaves_per_tree > 1, not the reference's algorithm): the simulations of this tree that are in flight below an edge count
// as visits that lost - N + vl visits, value (N * Q - vl) / (N + vl) - and the node's visit count includes them, which
// steers the next simulation of the same wave to another leaf.
// ------------------------------------------------------------------------------------------
template <bool VL, class W>
SPL_D int mcts_pick(const W& w, const MctsEdges& ed, const MctsEdgeV& first, int k, int Ns, float Qs, const MctsSearchParams& P, bool forced,
                    int n_iter) {   // `first` = edge [lane], fetched by the caller together with the node header
    const double fpu_init = P.fpu > 0.0 ? MC_DADD((double)Qs, -P.fpu) : P.fpu;   // :202
    double sq_ns, sq_ns_eps;
    w.sqrt2((double)Ns, MC_DADD((double)Ns, MCTS_EPS), sq_ns, sq_ns_eps);
    double best_u = 0.0;
    int best_i = -1, forced_i = 0x7fffffff;
    for (int i = w.lane; i < k; i += W::W) {
        MctsEdgeV e = i == w.lane ? first : ed.load(i);
        if (VL) {
            const int vl = (int)((uint32_t)e.N >> 24), n = e.N & MCTS_N_MASK;
            if (vl) {
                e.Q = MC_DDIV(MC_DADD(n ? MC_DMUL((double)n, e.Q) : 0.0, -(double)vl), (double)(n + vl));
                e.N = n + vl;
            } else e.N = n;
        }
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
template <int N, bool VL, class W>
SPL_D void mcts_backup(const W& w, const MctsArena& A, int t, int s, int depth, const float* v) {
    const uint32_t* path = mcts_path(A, t, s);
    for (int d = w.lane; d < depth; d += W::W) {
        const float v0 = mcts_vsel<N>(v, ((d - depth) % N + N) % N);
        const uint32_t rec = path[2 * d], pe = path[2 * d + 1];
        MctsNode* nd = mcts_node(A, rec);
        const MctsEdges ed = mcts_edges(A, rec, (int)(pe >> 16));
        const int pos = (int)(pe & 0xFFFFu);
        const MctsPN pn = ed.pn[pos];
        const int n = VL ? (pn.N & MCTS_N_MASK) : pn.N;
        ed.Q[pos] = MC_DDIV(MC_DADD(MC_DMUL((double)n, ed.Q[pos]), (double)v0), (double)(n + 1));                            // :171
        nd->u.x.Qs = MC_FDIV(MC_FADD(MC_FMUL((float)(nd->u.x.Ns + 1), nd->u.x.Qs), v0), (float)(nd->u.x.Ns + 2));          // :172
        ed.pn[pos].N = VL ? pn.N - MCTS_VL_ONE + 1 : pn.N + 1;      // (VL: the simulation is no longer in flight below this edge)
        nd->u.x.Ns += 1;
        if (VL) nd->u.x.pad -= 1u;
    }
    if (w.lane == 0) {
        MctsTree* T = A.trees + t;
#pragma unroll
        for (int i = 0; i < 4; i++) T->last_v[i] = i < N ? mcts_vsel<N>(v, ((i - depth) % N + N) % N) : 0.f;
        T->depth_sum += depth;
    }
}

// VL: a simulation that cannot go on this wave (another simulation of the tree already waits at the same node or edge) is
// given up: its virtual visits are taken back along the recorded path, nothing else was changed
#ifdef __CUDACC__
template <class W>
static __device__ __noinline__ void mcts_abandon(const W& w, const MctsArena& A, int t, int s, int depth) {
#else
template <class W>
SPL_D void mcts_abandon(const W& w, const MctsArena& A, int t, int s, int depth) {
#endif
    const uint32_t* path = mcts_path(A, t, s);
    for (int d = w.lane; d < depth; d += W::W) {
        const uint32_t rec = path[2 * d], pe = path[2 * d + 1];
        const MctsEdges ed = mcts_edges(A, rec, (int)(pe >> 16));
        ed.pn[pe & 0xFFFFu].N -= MCTS_VL_ONE;
        mcts_node(A, rec)->u.x.pad -= 1u;
    }
    if (w.lane == 0) {
        MctsSlot* S = A.trees[t].slot + s;
        S->leaf = 0u; S->cur = 0u; S->pend_edge = -1; S->path_len = 0;
    }
    w.sync();
}

// hands a node that waits for the network to the leaf row of its tree (:136-138)
#ifdef __CUDACC__
template <int N, class W>
static __device__ __noinline__ void mcts_emit_leaf(
#else
template <int N, class W>
SPL_D void mcts_emit_leaf(
#endif
const W& w, const MctsArena& A, int t, int s, uint32_t node, int depth, int sims_done, int8_t* leaf_state, uint8_t* leaf_valid,
                          bool rows = true) {   // rows = false: the staging row of the rules kernel already holds this very state and mask
    MctsTree* T = A.trees + t;
    if (rows) {
        const MctsNode* nd = mcts_node(A, node);
        mcts_decode<N>(w, mcts_cstate(A, node), leaf_state);
        for (int i = w.lane; i < SPL_ACTIONS; i += W::W) leaf_valid[i] = 0;
        w.sync();
        const int k = (int)nd->n_edges;
        const MctsEdges ed = mcts_edges(A, node, k);
        for (int i = w.lane; i < k; i += W::W) leaf_valid[ed.ca[i].action] = 1;
        if (w.lane == 0) A.leaf_src[mcts_row(A, t, s)] = 0;
    }
    if (w.lane == 0) {
        MctsSlot* S = T->slot + s;
        S->leaf = node; S->path_len = depth; S->cur = 0u; S->pend_edge = -1;
        if (sims_done >= 0) T->sims_done = sims_done;
    }
    w.sync();
}

// ------------------------------------------------------------------------------------------
// descent (light: no rules code): runs simulations of tree t (in-flight slot s) until one
//   * reaches a node that waits for the network  -> leaf row written, returns 1
//   * reaches an edge that was never taken       -> records it as pending (the rules kernel computes the child state,
//                                                    mcts_attach_tree links it), returns 2
//   * or the move's budget is spent              -> returns 0
// Simulations that end in a terminal node are backed up on the spot, at most `max_terminal` of them per call (then 3 is
// returned and the next call carries on): near the end of a game almost every simulation is such a one, and a single
// tree working through its whole budget would hold up the wave. For the same reason a call walks at most `max_levels`
// edges before it yields (3). Slots with a leaf or a pending edge are skipped.
// VL (leaves_per_tree > 1): the edges walked carry a virtual visit until the simulation is backed up; a simulation that runs
// into a node or an edge another slot of the tree already waits at is given up for this wave (returns 0).
// ------------------------------------------------------------------------------------------
template <int N, bool VL, class W>
SPL_D int mcts_descend_tree_impl(const W& w, const MctsArena& A, int t, int s, const MctsSearchParams& P, int max_terminal, int max_levels,
                                 int8_t* leaf_state, uint8_t* leaf_valid, int& spec_hits);
template <int N, bool VL, class W>
SPL_D int mcts_descend_tree(const W& w, const MctsArena& A, int t, int s, const MctsSearchParams& P, int max_terminal, int max_levels,
                            int8_t* leaf_state, uint8_t* leaf_valid) {
    int spec_hits = 0;   // diagnostics (early-fetch hit rate): counted in a register, written once per call - the tree belongs to this warp
    const int r = mcts_descend_tree_impl<N, VL>(w, A, t, s, P, max_terminal, max_levels, leaf_state, leaf_valid, spec_hits);
    if (spec_hits != 0 && w.lane == 0) A.trees[t].spec_hits += spec_hits;
    return r;
}
template <int N, bool VL, class W>
SPL_D int mcts_descend_tree_impl(const W& w, const MctsArena& A, int t, int s, const MctsSearchParams& P, int max_terminal, int max_levels,
                                 int8_t* leaf_state, uint8_t* leaf_valid, int& spec_hits) {
    MctsTree* T = A.trees + t;
    MctsSlot* S = T->slot + s;
    if (S->leaf != 0u) return 1;
    if (S->pend_edge >= 0) return 2;
    if (T->status != 0u) return 0;
    const uint32_t root = T->root;
    int sims_done = T->sims_done;
    int target = T->sims_target;
    uint32_t cur = S->cur;
    if (VL && cur == 0u) {   // a new simulation only if the budget is not already covered by the ones in flight
        int in_flight = 0;
        for (int j = 0; j < A.n_slots; j++)
            in_flight += (j != s && (T->slot[j].leaf != 0u || T->slot[j].pend_edge >= 0 || T->slot[j].cur != 0u)) ? 1 : 0;
        target -= in_flight;
    }
    if (root == 0u || (cur == 0u && sims_done >= target)) return 0;
    uint32_t* path = mcts_path(A, t, s);
    const uint32_t flags = T->flags;
    int depth = cur != 0u ? S->path_len : 0;
    if (cur == 0u) cur = root;
    // the edge count of `cur`: known from the parent's edge, or from the node header at the start of a walk
    int ne = mcts_node(A, cur)->kind == MCTS_NODE_TERMINAL ? 0 : (int)mcts_node(A, cur)->n_edges;
    // The child the PREVIOUS visit of a node chose is fetched (header + edges) while this visit is still computing its pick: the
    // search tends to walk the same line again, and then the next level starts without a memory round trip. The hint lives in
    // a spare word of the node header (edge position + 1); it never influences a result, only what is loaded early.
    bool have_spec = false;
    MctsEdgeV spec_first = mcts_edge_none();
    MctsNode spec_nd;
    for (;;) {
        for (;;) {
            const MctsEdges ed = mcts_edges(A, cur, ne);
            MctsEdgeV first;                                  // header and first 32 edges: one round trip
            MctsNode nd;
            if (have_spec) {
                first = spec_first; nd = spec_nd;
            } else {
                first = mcts_edge_none();
                if (w.lane < ne) first = ed.load(w.lane);
                nd = *mcts_node(A, cur);
            }
            have_spec = false;
            const int kind = nd.kind;
            if (kind == MCTS_NODE_NEEDS_NN) {
                if (VL) {
                    bool taken = false;
                    for (int j = 0; j < A.n_slots; j++) taken |= j != s && T->slot[j].leaf == cur;
                    if (taken) { mcts_abandon(w, A, t, s, depth); return 0; }
                }
                mcts_emit_leaf<N>(w, A, t, s, cur, depth, VL ? -1 : sims_done, leaf_state, leaf_valid);
                return 1;
            }
            if (kind == MCTS_NODE_TERMINAL) break;   // :130-132
            int hp = -1;
            const int hint = (int)nd.u.x.hint - 1;
#ifdef __CUDACC__
            if (!VL && hint >= 0 && hint < 32 && hint < (int)nd.n_edges) {
                const uint32_t hchild = __shfl_sync(0xffffffffu, first.child, hint);
                const int hne = __shfl_sync(0xffffffffu, (int)first.child_ne, hint);
                if (hchild != 0u) {
                    hp = hint;
                    spec_first = mcts_edge_none();
                    if (w.lane < hne) spec_first = mcts_edges(A, hchild, hne).load(w.lane);
                    spec_nd = *mcts_node(A, hchild);
                }
            }
            long long prof_t0 = 0;
            if (A.prof) prof_t0 = clock64();
#endif
            const int ns_eff = VL ? nd.u.x.Ns + (int)nd.u.x.pad : nd.u.x.Ns;
            const int ei = mcts_pick<VL>(w, ed, first, (int)nd.n_edges, ns_eff, nd.u.x.Qs, P, depth == 0 && (flags & MCTS_F_FORCED), sims_done);
#ifdef __CUDACC__
            if (A.prof && w.lane == 0) { A.prof[(size_t)t * 16 + 6] += clock64() - prof_t0; A.prof[(size_t)t * 16 + 7] += 1; }   // diagnostics: cycles inside the pick, levels
#endif
            uint32_t child;
            int child_ne;
#ifdef __CUDACC__
            if (ei < 32) {
                child = __shfl_sync(0xffffffffu, first.child, ei);
                child_ne = __shfl_sync(0xffffffffu, (int)first.child_ne, ei);
            } else
#endif
            {
                const MctsCA sel = ed.ca[ei];
                child = sel.child; child_ne = (int)sel.child_ne;
            }
            if (VL && child == 0u) {   // an edge another simulation of this wave is already opening
                bool taken = false;
                for (int j = 0; j < A.n_slots; j++)
                    taken |= j != s && T->slot[j].pend_edge >= 0 && T->slot[j].pend_parent == cur && (T->slot[j].pend_edge & 0xFFFF) == ei;
                if (taken) { mcts_abandon(w, A, t, s, depth); return 0; }
            }
            if (w.lane == 0) {
                path[2 * depth] = cur; path[2 * depth + 1] = (uint32_t)ei | ((uint32_t)nd.n_edges << 16);
                if (ei != hint) mcts_node(A, cur)->u.x.hint = (uint32_t)ei + 1u;
                if (VL) { ed.pn[ei].N += MCTS_VL_ONE; mcts_node(A, cur)->u.x.pad += 1u; }
            }
            depth++;
            if (child != 0u && --max_levels <= 0) {   // yield: a very deep path finishes in the next call instead of holding up the wave
                if (w.lane == 0) { S->cur = child; S->path_len = depth; if (!VL) T->sims_done = sims_done; }
                w.sync();
                return 3;
            }
            if (child == 0u) {   // first traversal of this edge
                if (w.lane == 0) {
                    S->pend_edge = (int32_t)((uint32_t)ei | ((uint32_t)nd.n_edges << 16)); S->pend_parent = cur;
                    S->path_len = depth; S->cur = 0u;
                    if (!VL) T->sims_done = sims_done;
                }
                w.sync();
                return 2;
            }
            cur = child;
            ne = child_ne;
            have_spec = ei == hp;
            spec_hits += have_spec ? 1 : 0;
        }
        {   // terminal: return Es up the path
            float v[N];
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = mcts_node(A, cur)->u.es[i];
            w.sync();
            mcts_backup<N, VL>(w, A, t, s, depth, v);
        }
        w.sync();
        sims_done = VL ? T->sims_done + 1 : sims_done + 1;
        if (VL) {
            if (w.lane == 0) { T->sims_done = sims_done; S->cur = 0u; S->path_len = 0; }
            w.sync();
            return 3;   // one simulation per slot and call
        }
        cur = root; depth = 0;
        ne = mcts_node(A, root)->kind == MCTS_NODE_TERMINAL ? 0 : (int)mcts_node(A, root)->n_edges;
        if (sims_done >= target) break;
        if (--max_terminal <= 0) {
            if (w.lane == 0) { T->sims_done = sims_done; S->cur = 0u; S->path_len = 0; }
            w.sync();
            return 3;
        }
    }
    if (w.lane == 0) { T->sims_done = sims_done; S->cur = 0u; S->path_len = 0; }
    w.sync();
    return 0;
}

// attaches the child state computed for the pending edge: st = its bytes (sp, zero padded, 16-aligned), ended / es /
// mask from mcts_rules_core. The dictionary may already hold the state (transposition, :120): then the descent goes on
// from that node in the next mcts_descend_tree call. Returns 1 if a leaf row was written, 0 otherwise.
template <int N, bool VL, class W>
SPL_D int mcts_attach_tree(const W& w, const MctsArena& A, int t, int s, const MctsSearchParams& P, const int8_t* st_aos, uint8_t* st /* cp bytes of scratch */,
                           bool ended, const float* es, const uint32_t* m, int mstride, int8_t* leaf_state, uint8_t* leaf_valid, bool emit_rows = true) {
    MctsTree* T = A.trees + t;
    MctsSlot* S = T->slot + s;
    const int pe = S->pend_edge;
    if (pe < 0) return S->leaf != 0u ? 1 : 0;
    uint32_t rec = 0u;
    if (mcts_encode<N, false>(w, st_aos, st)) {   // (the child of a node of this tree: made by the rules code from an accepted state)
        const uint64_t h = mcts_hash(w, st, A.cp);
        rec = mcts_lookup(w, A, t, st, h);
        if (rec == 0u) rec = mcts_store_node<N>(w, A, t, st, h, ended, es, m, mstride);
    } else if (w.lane == 0) T->status |= MCTS_S_BAD_STATE;   // cannot happen for states the rules produced from a representable root
    w.sync();
    if (rec == 0u) {   // no room: the search of this tree stops here (status bit set)
        if (VL) mcts_abandon(w, A, t, s, S->path_len);
        if (w.lane == 0) { S->pend_edge = -1; S->cur = 0u; S->path_len = 0; }
        w.sync();
        return 0;
    }
    const MctsNode* cn = mcts_node(A, rec);
    const int kind = cn->kind;
    if (w.lane == 0) {
        MctsCA* e = mcts_edges(A, S->pend_parent, pe >> 16).ca + (pe & 0xFFFF);
        e->child = rec;
        e->child_ne = kind == MCTS_NODE_TERMINAL ? (uint16_t)0 : cn->n_edges;
    }
    w.sync();
    const int depth = S->path_len, sims_done = T->sims_done;
    if (kind == MCTS_NODE_NEEDS_NN) {
        if (VL) {   // a transposition into a node another simulation of this wave already waits at
            bool taken = false;
            for (int j = 0; j < A.n_slots; j++) taken |= j != s && T->slot[j].leaf == rec;
            if (taken) { mcts_abandon(w, A, t, s, depth); return 0; }
        }
        mcts_emit_leaf<N>(w, A, t, s, rec, depth, VL ? -1 : sims_done, leaf_state, leaf_valid, emit_rows);
        return 1;
    }
    if (kind == MCTS_NODE_TERMINAL) {
        float v[N];
#pragma unroll
        for (int i = 0; i < N; i++) v[i] = cn->u.es[i];
        mcts_backup<N, VL>(w, A, t, s, depth, v);
        if (w.lane == 0) { T->sims_done = sims_done + 1; S->pend_edge = -1; S->cur = 0u; S->path_len = 0; }
    } else {   // an expanded node reached through a new edge: keep descending from it
        if (w.lane == 0) { S->pend_edge = -1; S->cur = rec; }
    }
    w.sync();
    return 0;
}

// expansion + backup: pi = the network's probability row for the leaf (float32[406], masked softmax), v = float32[N]
template <int N, bool VL, class W>
SPL_D void mcts_expand_tree(const W& w, const MctsArena& A, int t, int s, const MctsSearchParams& P, const float* pi, const float* vin,
                            const double* dir, double* dscratch) {
    MctsTree* T = A.trees + t;
    MctsSlot* S = T->slot + s;
    const uint32_t leaf = S->leaf;
    if (leaf == 0u) return;
    MctsNode* nd = mcts_node(A, leaf);
#ifdef __CUDACC__
    if (!VL) {
        // The common case in three rounds of loads instead of seven dependent ones: (1) the leaf's header, the value vector and
        // the recorded path, (2) the leaf's edge actions and the statistics of the path's nodes and edges, (3) the network's
        // probabilities; the float32 normalisation sums in action order through shuffles (same additions as `normalise :239`),
        // the backup (:168-177) uses the statistics fetched in round 2. Same arithmetic as the general path below.
        const int depth = S->path_len;
        const bool noise = leaf == T->root && T->sims_done == 0 && (T->flags & MCTS_F_NOISE);
        const MctsNode hdr = *nd;
        const int k = hdr.n_edges;
        if (!noise && k <= 32 && depth <= 32) {
            const uint32_t* path = mcts_path(A, t, s);
            float v[N];
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = vin[i];
            uint32_t pn = 0u, pe = 0u;
            if (w.lane < depth) { pn = path[2 * w.lane]; pe = path[2 * w.lane + 1]; }
            const MctsEdges ed = mcts_edges(A, leaf, k);
            int act = 0;
            if (w.lane < k) act = ed.ca[w.lane].action;
            int eN = 0, nNs = 0;
            double eQ = 0.0;
            float nQs = 0.f;
            MctsNode* pnode = mcts_node(A, pn);
            const MctsEdges ped = mcts_edges(A, pn, (int)(pe >> 16));
            const int ppos = (int)(pe & 0xFFFFu);
            if (w.lane < depth) { eN = ped.pn[ppos].N; eQ = ped.Q[ppos]; nNs = pnode->u.x.Ns; nQs = pnode->u.x.Qs; }
            float p = w.lane < k ? pi[act] : 0.f;
            float sm = 0.f;
            for (int i = 0; i < k; i++) sm = MC_FADD(sm, __shfl_sync(0xffffffffu, p, i));
            if (w.lane < k) ed.pn[w.lane].P = MC_FDIV(p, sm);
            if (w.lane == 0) {
                nd->u.x.Ns = 0;
                nd->u.x.Qs = v[0];   // :147
                nd->kind = MCTS_NODE_EXPANDED;
            }
            if (w.lane < depth) {
                const float v0 = mcts_vsel<N>(v, ((w.lane - depth) % N + N) % N);
                ped.Q[ppos] = MC_DDIV(MC_DADD(MC_DMUL((double)eN, eQ), (double)v0), (double)(eN + 1));                 // :171
                pnode->u.x.Qs = MC_FDIV(MC_FADD(MC_FMUL((float)(nNs + 1), nQs), v0), (float)(nNs + 2));               // :172
                ped.pn[ppos].N = eN + 1;
                pnode->u.x.Ns = nNs + 1;
            }
            if (w.lane == 0) {
#pragma unroll
                for (int i = 0; i < 4; i++) T->last_v[i] = i < N ? mcts_vsel<N>(v, ((i - depth) % N + N) % N) : 0.f;
                T->depth_sum += depth;
                T->sims_done += 1;
                S->leaf = 0u; S->cur = 0u; S->pend_edge = -1; S->path_len = 0;
                T->nn_calls += 1;
            }
            w.sync();
            return;
        }
    }
#endif
    const int k = nd->n_edges;
    const MctsEdges ed = mcts_edges(A, leaf, k);
    for (int i = w.lane; i < k; i += W::W) ed.pn[i].P = pi[ed.ca[i].action];
    w.sync();
    if (leaf == T->root && T->sims_done == 0 && (T->flags & MCTS_F_NOISE)) {   // :141-143
        mcts_root_noise(w, ed.pn, k, P, dir, P.game_base + (uint32_t)t, T->episode, (uint32_t)nd->ply, dscratch);
    } else {                                                                      // normalise :144
        if (w.lane == 0) {
            float sm = 0.f;
            for (int i = 0; i < k; i++) sm = MC_FADD(sm, ed.pn[i].P);
            reinterpret_cast<float*>(dscratch)[0] = sm;
        }
        w.sync();
        const float sm = reinterpret_cast<float*>(dscratch)[0];
        for (int i = w.lane; i < k; i += W::W) ed.pn[i].P = MC_FDIV(ed.pn[i].P, sm);
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
    mcts_backup<N, VL>(w, A, t, s, S->path_len, v);
    w.sync();
    if (w.lane == 0) {
        T->sims_done += 1;
        S->leaf = 0u; S->cur = 0u; S->pend_edge = -1; S->path_len = 0;
        T->nn_calls += 1;
    }
    w.sync();
}

// ------------------------------------------------------------------------------------------
// cleaning. What the reference's dictionary holds but no later search of the same game can look up again:
//   * nodes below the root's ply (every key carries its ply; the reference's own cleaning :80-85 relies on the same fact), and
//   * nodes whose deck rows differ from the root's: moves inside the tree are deterministic (make_move(..., deterministic=True),
//     :228) and never touch the decks (SplendorLogicNumba.py:445-450,529-532), so every state a search from this root or from
//     any later root of the game looks up carries a deck that is a subset of the root's; a node with a card in its deck
//     that the root's deck no longer has can never be equal to one of them.
// Dropping those is result-neutral ("exact"). A real move that reveals a card therefore retires the whole tree at the next
// begin in O(pages) - no copying -, and within a line of moves without reveals the tree simply grows in the shared pool.
// When a tree reaches its node limit the survivors are COPIED into fresh pages (every copy independent: lane per record,
// forwarding pointers in the old headers), the hash table is rebuilt and the old pages go back to the pool.
// ------------------------------------------------------------------------------------------
SPL_D void mcts_deck_of(const uint8_t* cst, uint8_t* deck15) {   // of a compact state
    for (int i = 0; i < 15; i++) deck15[i] = cst[MCTS_C_DECK + i];
}
SPL_D bool mcts_deck_subset(const uint8_t* cst, const uint8_t* deck15) {   // every card still in the deck of `cst` is still in deck15
    bool ok = true;
    for (int i = 0; i < 15; i++) ok &= (cst[MCTS_C_DECK + i] & ~deck15[i]) == 0;
    return ok;
}

// gives every page of the tree back and empties its table; statistics and the simulation in flight are the caller's business
template <class W>
SPL_D void mcts_release_storage(const W& w, const MctsArena& A, int t) {
    MctsTree* T = A.trees + t;
    uint32_t* tab = A.htab + (size_t)t * A.hcap;
    mcts_pages_push(w, A, A.tree_pages + (size_t)t * A.max_pages, T->n_pages);
    uint4* t4 = reinterpret_cast<uint4*>(tab);
    uint4 z; z.x = z.y = z.z = z.w = 0u;
    for (int i = w.lane; i < A.hcap / 4; i += W::W) t4[i] = z;
    w.sync();
    if (w.lane == 0) { T->n_pages = 0; T->bump = 0u; T->page_end = 0u; T->n_nodes = 0; T->n_edges = 0; }
    w.sync();
}

template <class W>
SPL_D void mcts_clear_tree(const W& w, const MctsArena& A, int t) {   // reset_all_search_trees :188-192 for one tree
    mcts_release_storage(w, A, t);
    if (w.lane == 0) {
        MctsTree* T = A.trees + t;
        T->root = 0u; T->sims_done = 0; T->sims_target = 0;
        T->flags = 0u; T->status = 0u; T->depth_sum = 0; T->spec_hits = 0;
        for (int j = 0; j < MCTS_KMAX; j++) { MctsSlot* S = T->slot + j; S->leaf = 0u; S->cur = 0u; S->pend_edge = -1; S->pend_parent = 0u; S->path_len = 0; }
        T->n_dropped = 0; T->hetero = 0;
    }
    w.sync();
}

// every node of the tree has become unreachable (the root's deck lost a card): an exact cleaning that copies nothing
template <class W>
SPL_D void mcts_retire_all(const W& w, const MctsArena& A, int t) {
    MctsTree* T = A.trees + t;
    const int n = T->n_nodes;
    mcts_release_storage(w, A, t);
    if (w.lane == 0) {
        T->n_dropped += n; T->compactions += 1;
        T->root = 0u; T->hetero = 0;
        for (int j = 0; j < MCTS_KMAX; j++) { MctsSlot* S = T->slot + j; S->leaf = 0u; S->cur = 0u; S->pend_edge = -1; S->pend_parent = 0u; S->path_len = 0; }
    }
    w.sync();
}

// marks (fwd = 1) every node reachable from `root` through linked edges: a stack threaded through the fwd words of the
// records themselves, one node popped at a time, its edges looked at by all lanes
template <class W>
SPL_D void mcts_mark_reachable(const W& w, const MctsArena& A, uint32_t root) {
    if (w.lane == 0) mcts_node(A, root)->fwd = 1u;   // bottom of the stack: link = 1 = "marked, nothing below"
    w.sync();
    uint32_t head = root;
    while (head != 1u) {
        const uint32_t rec = head;
        MctsNode* nd = mcts_node(A, rec);
        head = nd->fwd;
        const int k = nd->kind == MCTS_NODE_TERMINAL ? 0 : (int)nd->n_edges;
        w.sync();
        if (w.lane == 0) nd->fwd = 1u;
        w.sync();
        const MctsEdges ed = mcts_edges(A, rec, k);
        for (int base = 0; base < k; base += W::W) {
            const int i = base + w.lane;
            uint32_t c = i < k ? ed.ca[i].child : 0u;
            bool fresh = false;
            if (c) {
#ifdef __CUDACC__
                fresh = atomicCAS(&mcts_node(A, c)->fwd, 0u, 1u) == 0u;   // two edges of a node may lead to the same child
#else
                fresh = mcts_node(A, c)->fwd == 0u;
#endif
            }
            // the fresh children go on the stack in lane order: each links to the previous fresh one, the first to the old head
            const uint32_t b = w.ballot(fresh);
#ifdef __CUDACC__
            if (b) {
                const int prev_lane = 31 - __clz((int)(b & w.lanemask_lt()));       // -1 if none below
                const uint32_t prev_c = __shfl_sync(0xffffffffu, c, prev_lane < 0 ? 0 : prev_lane);
                if (fresh) mcts_node(A, c)->fwd = prev_lane < 0 ? head : prev_c;
                head = __shfl_sync(0xffffffffu, c, 31 - __clz((int)b));
            }
#else
            if (b) { mcts_node(A, c)->fwd = head; head = c; }
#endif
            w.sync();
        }
    }
    w.sync();
}

// Copies the surviving records of tree t into fresh pages. use_marks: survivors = records with fwd != 0 (mcts_mark_reachable);
// else the exact rule: ply >= min_ply and (check_deck: deck within deck15). The simulation in flight (root, cur, leaf, pending
// edge, recorded path) is re-based. Returns false, with the tree unchanged, when the pool cannot supply the pages.
template <class W>
SPL_D bool mcts_compact(const W& w, const MctsArena& A, int t, bool use_marks, int min_ply, const uint8_t* deck15, bool check_deck) {
    MctsTree* T = A.trees + t;
    uint32_t* tab = A.htab + (size_t)t * A.hcap;
    uint32_t* pages = A.tree_pages + (size_t)t * A.max_pages;
    const int old_pages = T->n_pages, n_old = T->n_nodes;
    const uint32_t old_bump = T->bump, old_end = T->page_end;
    if (w.lane == 0) { T->bump = 0u; T->page_end = 0u; }   // the copy starts on a page of its own
    w.sync();
    uint32_t chain = 0u;   // the new records this lane made, linked through their fwd words
    int n_live = 0, e_live = 0;
    bool fail = false;
    for (int base = 0; base < A.hcap && !fail; base += W::W) {
        const int s = base + w.lane;
        const uint32_t rec = s < A.hcap ? tab[s] : 0u;
        bool live = false;
        int k = 0;
        if (rec) {
            const MctsNode* nd = mcts_node(A, rec);
            k = nd->kind == MCTS_NODE_TERMINAL ? 0 : (int)nd->n_edges;
            if (use_marks) live = nd->fwd != 0u;
            else live = (int)nd->ply >= min_ply && (!check_deck || mcts_deck_subset(mcts_cstate(A, rec), deck15));
        }
        const uint32_t lb = w.ballot(live);
        if (lb == 0u) continue;
        const uint32_t units = live ? mcts_rec_units(A, k) : 0u;
        int total = 0;
        const int excl = w.scan_excl((int)units, total);
        uint32_t dst = 0u;
        uint32_t cur_bump = T->bump, cur_end = T->page_end;
        if (cur_bump != 0u && cur_bump + (uint32_t)total <= cur_end) {   // the whole batch fits on the current page
            dst = cur_bump + (uint32_t)excl;
            w.sync();
            if (w.lane == 0) T->bump = cur_bump + (uint32_t)total;
            w.sync();
        } else {                                                       // page boundary: one record at a time, in lane order
            for (int l = 0; l < W::W; l++) {
                if (!((lb >> l) & 1u)) continue;
                const uint32_t u = (uint32_t)w.shfl((int)units, l);
                const uint32_t o = mcts_alloc(w, A, t, u, true);
                if (o == 0u) { fail = true; break; }
                if (w.lane == l) dst = o;
            }
        }
        if (fail) break;
        if (live) {
            const uint4* s4 = reinterpret_cast<const uint4*>(mcts_ptr(A, rec));
            uint4* d4 = reinterpret_cast<uint4*>(mcts_ptr(A, dst));
            const int n16 = (int)units * 2;
            int i = 0;
            for (; i + 4 <= n16; i += 4) {
                const uint4 a = s4[i], b = s4[i + 1], c = s4[i + 2], d = s4[i + 3];
                d4[i] = a; d4[i + 1] = b; d4[i + 2] = c; d4[i + 3] = d;
            }
            for (; i < n16; i++) d4[i] = s4[i];
            mcts_node(A, rec)->fwd = dst;
            mcts_node(A, dst)->fwd = chain;
            chain = dst;
            n_live += 1; e_live += k;
        }
    }
    fail = w.ballot(fail) != 0u;
    if (fail) {   // out of pages: give the copy back, forget the forwarding pointers
        w.sync();
        mcts_pages_push(w, A, pages + old_pages, T->n_pages - old_pages);
        for (int s = w.lane; s < A.hcap; s += W::W) {
            const uint32_t rec = tab[s];
            if (rec) mcts_node(A, rec)->fwd = 0u;
        }
        w.sync();
        if (w.lane == 0) { T->n_pages = old_pages; T->bump = old_bump; T->page_end = old_end; }
        w.sync();
        return false;
    }
    w.sync();
    // the simulations in flight follow their nodes (a dropped node has fwd == 0)
    for (int j = 0; j < A.n_slots; j++) {
        MctsSlot* S = T->slot + j;
        uint32_t* path = mcts_path(A, t, j);
        const bool in_flight = S->leaf != 0u || S->cur != 0u || S->pend_edge >= 0;
        const int plen = in_flight ? S->path_len : 0;
        for (int d = w.lane; d < plen; d += W::W) path[2 * d] = mcts_node(A, path[2 * d])->fwd;
        w.sync();
        if (w.lane == 0) {
            if (S->leaf) S->leaf = mcts_node(A, S->leaf)->fwd;
            if (S->cur) S->cur = mcts_node(A, S->cur)->fwd;
            if (S->pend_edge >= 0) S->pend_parent = mcts_node(A, S->pend_parent)->fwd;
        }
        w.sync();
    }
    if (w.lane == 0 && T->root) T->root = mcts_node(A, T->root)->fwd;     // (0 if the root itself was dropped: begin re-creates it)
    w.sync();
    // new table; every lane walks the records it copied: children follow the forwarding pointers of the old copies
    {
        uint4* t4 = reinterpret_cast<uint4*>(tab);
        uint4 z; z.x = z.y = z.z = z.w = 0u;
        for (int i = w.lane; i < A.hcap / 4; i += W::W) t4[i] = z;
    }
    w.sync();
    for (uint32_t rec = chain; rec != 0u;) {
        MctsNode* nd = mcts_node(A, rec);
        const uint32_t next = nd->fwd;
        nd->fwd = 0u;
        const int k = nd->kind == MCTS_NODE_TERMINAL ? 0 : (int)nd->n_edges;
        MctsCA* ca = mcts_edges(A, rec, k).ca;
        for (int e = 0; e < k; e++) {
            const uint32_t c = ca[e].child;
            if (c) ca[e].child = mcts_node(A, c)->fwd;
        }
        uint32_t slot = (uint32_t)nd->hash & (uint32_t)(A.hcap - 1);
        for (;;) {
#ifdef __CUDACC__
            if (atomicCAS(&tab[slot], 0u, rec) == 0u) break;
#else
            if (tab[slot] == 0u) { tab[slot] = rec; break; }
#endif
            slot = (slot + 1u) & (uint32_t)(A.hcap - 1);
        }
        rec = next;
    }
    w.sync();
    // the old pages go back to the pool; the new ones move to the front of the tree's page list
    mcts_pages_push(w, A, pages, old_pages);
    const int new_pages = T->n_pages - old_pages;
    for (int base = 0; base < new_pages; base += W::W) {
        const int i = base + w.lane;
        uint32_t p = 0u;
        if (i < new_pages) p = pages[old_pages + i];
        w.sync();
        if (i < new_pages) pages[i] = p;
        w.sync();
    }
    n_live = w.sum(n_live); e_live = w.sum(e_live);
    if (w.lane == 0) {
        T->n_pages = new_pages;
        if (!use_marks) T->n_dropped += n_old - n_live;
        T->n_nodes = n_live; T->n_edges = e_live;
        T->compactions += 1;
    }
    w.sync();
    return true;
}

// periodic cleaning between waves (any state of the search): trees that hold more than max_nodes nodes drop what can no
// longer be used - per the exact rule, or everything the root does not reach (gc_reachable)
template <class W>
SPL_D void mcts_clean_tree(const W& w, const MctsArena& A, int t, int max_nodes, int gc_reachable) {
    MctsTree* T = A.trees + t;
    if (T->root == 0u || T->n_nodes <= max_nodes) return;
    if (gc_reachable) {
        mcts_mark_reachable(w, A, T->root);
        mcts_compact(w, A, t, true, 0, nullptr, false);
    } else {
        uint8_t deck[15];
        mcts_deck_of(mcts_cstate(A, T->root), deck);
        mcts_compact(w, A, t, false, (int)mcts_node(A, T->root)->ply, deck, T->hetero != 0);
    }
}

// root_state: the reference's int8[R,7] bytes.
// gc_reachable = 0: cleaning is exact (see above); gc_reachable = 1: a tree that reaches its node limit keeps only what the new
// root reaches (a node that only a not-yet-linked edge transposes into is re-created instead of found, which the reference's
// dictionary would not do).
template <int N, class W>
SPL_D void mcts_begin_tree(const W& w, const MctsArena& A, int t, const MctsSearchParams& P, const int8_t* root_state, int sims_target,
                           uint32_t flags, int gc_reachable, const double* dir, uint32_t episode, int8_t* st, uint8_t* cst /* cp bytes */, uint32_t* scratch,
                           double* dscratch) {
    typedef MctsLay<N> ML;
    MctsTree* T = A.trees + t;
    for (int i = w.lane; i < ML::SP; i += W::W) st[i] = i < ML::S ? root_state[i] : (int8_t)0;
    w.sync();
    if (!mcts_encode<N, true>(w, st, cst)) {   // a board the reference's rules cannot have produced (a row that is no card, a count that is not its mask's)
        if (w.lane == 0) { T->status |= MCTS_S_BAD_STATE; T->root = 0u; T->sims_done = 0; T->sims_target = 0; }
        w.sync();
        return;
    }
    if (A.n_slots > 1)      // simulations a truncated search left in flight give their virtual visits back
        for (int j = 0; j < A.n_slots; j++) {
            const MctsSlot* S = T->slot + j;
            if (S->leaf != 0u || S->cur != 0u || S->pend_edge >= 0) mcts_abandon(w, A, t, j, S->path_len);
        }
    const int root_ply = (int)(uint8_t)st[6];
    const uint64_t h = mcts_hash(w, cst, A.cp);
    uint8_t deck[15];
    mcts_deck_of(cst, deck);
    if (T->n_nodes > 0 && !T->hetero) {
        bool same = true, subset = true;
        for (int i = 0; i < 15; i++) { same &= deck[i] == T->deck[i]; subset &= (deck[i] & ~T->deck[i]) == 0; }
        if (!same) {
            if (subset) mcts_retire_all(w, A, t);          // a card was revealed since the tree's nodes were made
            else { if (w.lane == 0) T->hetero = 1; w.sync(); }   // not a later position of the same game: keep everything
        }
    }
    const int need_nodes = sims_target + 2;
    const uint32_t overflowed = T->status & (MCTS_S_OVERFLOW_NODES | MCTS_S_OVERFLOW_POOL);
    if (overflowed || T->n_nodes + need_nodes > A.node_limit) {
        if (gc_reachable) {
            const uint32_t old_root = mcts_lookup(w, A, t, cst, h);
            if (old_root) {
                mcts_mark_reachable(w, A, old_root);
                mcts_compact(w, A, t, true, 0, nullptr, false);
            } else {   // a root the tree has never seen: nothing of the old tree can be reached
                const int resets = T->resets;
                mcts_clear_tree(w, A, t);
                if (w.lane == 0) T->resets = resets;
                w.sync();
            }
        } else {
            mcts_compact(w, A, t, false, root_ply, deck, T->hetero != 0);
            if (T->n_nodes + need_nodes > A.node_limit) {
                // the exact cleaning did not free enough: fall back to the reachable set of the new root (counted as lossy:
                // nodes that only a not-yet-linked edge could transpose into are dropped)
                const uint32_t old_root = mcts_lookup(w, A, t, cst, h);
                if (old_root) {
                    mcts_mark_reachable(w, A, old_root);
                    mcts_compact(w, A, t, true, 0, nullptr, false);
                    if (w.lane == 0) T->resets += 1;
                    w.sync();
                }
            }
        }
        if (T->n_nodes + need_nodes > A.node_limit) {   // still no room: forget the tree (counted)
            const int resets = T->resets;
            mcts_clear_tree(w, A, t);
            if (w.lane == 0) T->resets = resets + 1;
            w.sync();
        }
        if (w.lane == 0) {
            if (overflowed) T->truncated += 1;
            T->status &= ~(uint32_t)(MCTS_S_OVERFLOW_NODES | MCTS_S_OVERFLOW_POOL | MCTS_S_BAD_STATE);
        }
        w.sync();
    }
    if (T->n_nodes == 0) {   // an empty tree takes the deck of its first root
        if (w.lane == 0) { for (int i = 0; i < 15; i++) T->deck[i] = deck[i]; T->hetero = 0; }
        w.sync();
    }
    uint32_t rec = mcts_lookup(w, A, t, cst, h);
    if (rec == 0u) rec = mcts_create_node<N>(w, A, t, P, st, cst, h, scratch);
    if (w.lane == 0) {
        T->root = rec; T->sims_done = 0; T->sims_target = rec == 0u ? 0 : sims_target;
        T->flags = flags; T->episode = episode;
        for (int j = 0; j < MCTS_KMAX; j++) { MctsSlot* S = T->slot + j; S->leaf = 0u; S->cur = 0u; S->pend_edge = -1; S->pend_parent = 0u; S->path_len = 0; }
    }
    w.sync();
    // a root the tree already expanded gets the noise on its stored Ps before the first simulation picks (:150-154);
    // a new root gets it when its network row arrives (mcts_expand_tree)
    if (rec != 0u && sims_target > 0 && (flags & MCTS_F_NOISE)) {
        const MctsNode* nd = mcts_node(A, rec);
        if (nd->kind == MCTS_NODE_EXPANDED)
            mcts_root_noise(w, mcts_edges(A, rec, (int)nd->n_edges).pn, (int)nd->n_edges, P, dir, P.game_base + (uint32_t)t, episode, (uint32_t)nd->ply, dscratch);
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
    if (T->root == 0u) return;
    const MctsNode* nd = mcts_node(A, T->root);
    if (nd->kind != MCTS_NODE_EXPANDED) return;
    const int k = nd->n_edges;
    const MctsEdges ed = mcts_edges(A, T->root, k);
    const bool forced = (T->flags & MCTS_F_FORCED) != 0u;
    if (w.lane == 0) {
        const float qs = nd->u.x.Qs;
        for (int p = 0; p < N; p++) q[p] = p == 0 ? (double)qs : (double)MC_FDIV(-qs, (float)(N - 1));   // :65-66
        int best = 0;
        for (int i = 0; i < k; i++) best = ed.pn[i].N > best ? ed.pn[i].N : best;
        double sum = 0.0;
        int besti = -1;
        double bestc = -1.0;
        for (int i = 0; i < k; i++) {
            const MctsPN e = ed.pn[i];
            double c = (double)e.N;
            if (forced) {   // policy target pruning :69-74
                if (e.N != best) c = c - (double)(long long)MC_DSQRT(MC_DMUL(MC_DMUL(MCTS_KFORCED, (double)e.P), (double)T->sims_target));
                c = c > 1.0 ? c : 0.0;
            }
            if (c > bestc) { bestc = c; besti = i; }
            if (temp != 0.0) {
                c = temp == 1.0 ? c : pow(c, 1.0 / temp);   // :94
                sum = MC_DADD(sum, c);
            }
            probs[ed.ca[i].action] = c;
        }
        if (temp == 0.0) {   // :87-92
            for (int i = 0; i < k; i++) probs[ed.ca[i].action] = i == besti ? 1.0 : 0.0;
        } else {
            for (int i = 0; i < k; i++) probs[ed.ca[i].action] = MC_DDIV(probs[ed.ca[i].action], sum);   // :95-96
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
    const bool fin = T->root != 0u && (T->sims_done >= T->sims_target || T->status != 0u);
    *finished = fin;
    if (!fin) return -1;
    const MctsNode* nd = mcts_node(A, T->root);
    if (nd->kind != MCTS_NODE_EXPANDED) return -1;
    const int k = nd->n_edges;
    const MctsEdges ed = mcts_edges(A, T->root, k);
    bool forced = (T->flags & MCTS_F_FORCED) != 0u;
    int action = -1;
    if (w.lane == 0) {
        int best = 0;
        for (int i = 0; i < k; i++) best = ed.pn[i].N > best ? ed.pn[i].N : best;
        double sum = 0.0, bestc = -1.0;
        int besti = -1;
        for (int pass = 0; pass < 2; pass++) {      // pass 0: total weight, pass 1: the walk up to u * total
            double target = 0.0, acc = 0.0;
            if (pass == 1) {
                if (forced && !(sum > 0.0) && best > 0) {
                    // policy-target pruning left nothing (every visited move was visited once: `c > 1 else 0`, :73): the reference
                    // divides by zero here and its caller's np.random.choice raises; the engine falls back to the raw counts
                    forced = false; pass = -1; sum = 0.0; bestc = -1.0; besti = -1;
                    continue;
                }
                if (temp == 0.0 || !(sum > 0.0)) break;
                const SplPhilox r = spl_philox(P.seed, P.game_base + (uint32_t)t, episode, (uint32_t)nd->ply, 5);
                target = MC_DMUL(mcts_u01(r.v[0], r.v[1]), sum);
            }
            for (int i = 0; i < k; i++) {
                const MctsPN e = ed.pn[i];
                double c = (double)e.N;
                if (forced) {   // policy target pruning :69-74
                    if (e.N != best) c = c - (double)(long long)MC_DSQRT(MC_DMUL(MC_DMUL(MCTS_KFORCED, (double)e.P), (double)T->sims_target));
                    c = c > 1.0 ? c : 0.0;
                }
                if (pass == 0 && c > bestc) { bestc = c; besti = i; }
                if (temp != 0.0 && temp != 1.0) c = pow(c, 1.0 / temp);   // :94
                if (pass == 0) sum = MC_DADD(sum, c);
                else {
                    acc = MC_DADD(acc, c);
                    if (c > 0.0) action = (int)ed.ca[i].action;      // the last action with weight, should rounding leave acc <= target at the end
                    if (acc > target) break;
                }
            }
        }
        if (temp == 0.0 && besti >= 0 && bestc > 0.0) action = (int)ed.ca[besti].action;
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
    if (T->root != 0u) {
        const MctsNode* nd = mcts_node(A, T->root);
        if (nd->kind == MCTS_NODE_EXPANDED) {
            const int k = (int)nd->n_edges;
            const MctsEdges ed = mcts_edges(A, T->root, k);
            for (int i = w.lane; i < k; i += W::W) {
                const int a = ed.ca[i].action;
                if (nsa) nsa[a] = ed.pn[i].N;
                if (qsa) qsa[a] = ed.Q[i];
                if (ps) ps[a] = ed.pn[i].P;
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
        info[14] = T->n_dropped;     // nodes exact cleanings dropped: n_nodes + n_dropped = size of the reference's dictionary
        info[15] = T->n_pages;
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
