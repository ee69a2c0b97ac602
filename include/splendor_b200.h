/* splendor_b200.h - C ABI of the B200-native batched Splendor environment (libsplendor_b200.so).
 *
 * The reference (kuboyoo/alphazero-general-ori) has no FFI for this path: its boundary is the
 * duck-typed Python `Game` protocol (Game.py:14-155, implemented by SplendorGame.py:11-86) over a
 * Numba jitclass `Board` (SplendorLogicNumba.py:84-774). This library is what sits underneath our
 * Python mirror of that protocol (alphazero-general-ori_b200/game.py); every entry point names the
 * reference call(s) it replaces. INTEGRATION.md shows the ctypes stub a reference maintainer adds.
 *
 * Conventions
 *  - plain C types only; all buffers are caller-owned DEVICE pointers (e.g. torch tensors' data_ptr)
 *  - every call takes the CUDA stream to launch on (cudaStream_t passed as void*; NULL = default stream)
 *    and returns 0 or a negative SPL_E_* code; nothing throws across the ABI; spl_last_error() gives text
 *  - no global state: a spl_ctx carries (n_players, rules, device); contexts are thread-compatible
 *  - there is no CPU fallback: without a CUDA device spl_ctx_create fails with SPL_E_NOGPU
 *
 * Data layouts (DESIGN.md "data layout")
 *  - AoS state  : int8[L][R][7], R = 32 + 10n + n^2  -- byte-identical to the reference's state array
 *                 (SplendorLogicNumba.py:291-303); used only at the API boundary
 *  - lane tiles : int8[Lpad/32][7R][32], Lpad = L rounded up to 32 -- the resident form in HBM. Games are
 *                 grouped in tiles of 32 lanes (one warp); inside a tile the state is structure-of-arrays
 *                 (cell-major), so a tile is one contiguous 7R*32-byte block that a single TMA bulk copy
 *                 stages into shared memory and lane l reads byte [cell][l]. Called "planes" below.
 *  - mask planes: uint32[13][Lpad] -- the 406 legality flags, bit a of the mask = action a
 */
#ifndef SPLENDOR_B200_H
#define SPLENDOR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPL_ABI_VERSION 1
#define SPL_NUM_ACTIONS 406     /* action_size(), SplendorLogicNumba.py:29-36 with patch P2 */
#define SPL_MASK_WORDS32 13
#define SPL_LANE_TILE 32
#define SPL_MAX_SYMMETRIES 18   /* 1 + 9 + 2n, get_symmetries :349-395 */

#define SPL_OK 0
#define SPL_E_ARG (-1)          /* bad argument */
#define SPL_E_CUDA (-2)         /* a CUDA call failed; see spl_last_error() */
#define SPL_E_NOGPU (-3)        /* no CUDA device / driver */

/* rule switches = Board fields ENABLE_ACTION_RESERVE / ENABLE_ACTION_GIVEBACK (SplendorLogicNumba.py:96-97);
 * REFCOMPAT reproduces the reference's n>=3 quirks (noble stride 3 at :219/:345, int8(999) at :313) */
#define SPL_RULE_RESERVE 1u
#define SPL_RULE_GIVEBACK 2u
#define SPL_RULE_REFCOMPAT 4u
#define SPL_RULES_DEFAULT (SPL_RULE_RESERVE | SPL_RULE_GIVEBACK | SPL_RULE_REFCOMPAT)

/* chance modes of a move (make_move's `deterministic` flag, :267, plus the replay used for parity) */
#define SPL_CHANCE_DETERMINISTIC 0   /* no deck reveal: the in-tree MCTS step (MCTS.py:228) */
#define SPL_CHANCE_REPLAY 1          /* reveal the given card: colour*8+idx per lane, 255 = none */
#define SPL_CHANCE_PHILOX 2          /* Philox4x32-10 keyed by (seed, game, episode, ply) */

typedef struct spl_ctx spl_ctx;

int         spl_abi_version(void);
const char* spl_last_error(void);

/* Board.__init__ (:86-98): n_players in {2,3,4}; token_limit = NUM_TOKEN_LIMIT (10) */
int  spl_ctx_create(int n_players, int token_limit, uint32_t rule_flags, int device, spl_ctx** out);
/* setNumTokenLim (:214), SplendorGame.disableReserve/enableReserve (SplendorGame.py:82-86) */
int  spl_ctx_set_rules(spl_ctx* ctx, int token_limit, uint32_t rule_flags);
/* 1 (default): the step kernel stages tiles with TMA; 0: plain vectorised copies (A/B and debugging) */
int  spl_ctx_set_tma(spl_ctx* ctx, int enabled);
void spl_ctx_destroy(spl_ctx* ctx);

/* observation_size (:25-27) and derived sizes */
int    spl_state_rows(int n_players);
int    spl_state_bytes(int n_players);
int    spl_lanes_padded(int n_lanes);
size_t spl_planes_bytes(int n_players, int n_lanes);
size_t spl_mask_planes_bytes(int n_lanes);

/* AoS <-> lane planes (Board.copy_state :291-303 is the reference's "bind this array") */
int spl_pack(spl_ctx* ctx, const int8_t* aos, int8_t* planes, int n_lanes, void* stream);
int spl_unpack(spl_ctx* ctx, const int8_t* planes, int8_t* aos, int n_lanes, void* stream);
/* mask planes -> bool[L][406] as the reference's valid_moves returns (:251-265) */
int spl_mask_unpack(spl_ctx* ctx, const uint32_t* mask_planes, uint8_t* valids, int n_lanes, void* stream);

/* init_game (:222-246). Philox: 12 deals + n+1 nobles from (seed, game_base+lane, episode[lane]).
 * lane_select (may be NULL) restricts the reset to lanes with a non-zero byte. */
int spl_reset_philox(spl_ctx* ctx, int8_t* planes, int n_lanes, uint64_t seed, uint32_t game_base,
                     const uint32_t* episodes, const uint8_t* lane_select, void* stream);
/* explicit start for replaying a reference game: deals uint8[L][12] (colour*8+idx per visible slot),
 * nobles uint8[L][5] (first n+1 used) */
int spl_reset_explicit(spl_ctx* ctx, int8_t* planes, int n_lanes, const uint8_t* deals,
                       const uint8_t* nobles, void* stream);

/* One fused pass over every lane (one thread per game lane):
 *   make_move (:267-289) -> [swap_players (:338-347)] -> check_end_game (:320-334) -> [auto reset]
 *   -> valid_moves (:251-265) for the player to move -> [uniform random pick for the next ply]
 * i.e. SplendorGame.getNextState + getCanonicalForm + getGameEnded + getValidMoves
 * (SplendorGame.py:30-57) for all lanes in one launch. Any output pointer may be NULL. */
typedef struct {
    int8_t*        planes;        /* in/out lane planes */
    int            n_lanes;
    const int16_t* actions;       /* [L] action per lane; <0 or NULL: no move (query only) */
    const uint8_t* players;       /* [L] index of the mover / queried player; NULL: `player` for all */
    int            player;
    int            chance_mode;   /* SPL_CHANCE_* */
    const uint8_t* reveals;       /* [L] for SPL_CHANCE_REPLAY */
    uint64_t       seed;          /* Philox key */
    uint32_t       game_base;     /* game id of lane 0 (global lane index when sharded over GPUs) */
    uint32_t*      episodes;      /* [L] in/out episode counters (NULL: 0) */
    int            rotate;        /* 1: rotate so that the next player becomes player 0 (canonical form) */
    int            auto_reset;    /* 1 (needs rotate=1, episodes): finished lanes start a new Philox game */
    int            store_state;   /* 0: leave planes untouched (pure queries) */
    uint32_t*      mask_out;      /* mask planes for the player to move after this call */
    float*         ended_out;     /* [L][n] check_end_game of the state as stored (before any auto reset) */
    int16_t*       next_actions;  /* [L] uniform random legal action for the next ply (Philox stream 1) */
    int32_t*       status_out;    /* [L] next player in the caller's frame, or <0 if the move was refused */
    unsigned long long* counters; /* [2] += finished games, += plies played (may be NULL) */
} spl_step_args;

int spl_step(spl_ctx* ctx, const spl_step_args* args, void* stream);

/* Persistent multi-ply pass: every lane plays `plies` plies of uniformly random legal moves
 * (Philox stream 1) with Philox deck reveals, end-game detection and auto reset, the tile staying in
 * shared memory for the whole call. This is the loop the reference runs one Python call at a time in
 * Arena.playGame (Arena.py:99-160) with SplendorPlayers.RandomPlayer (SplendorPlayers.py:20-27):
 * getValidMoves -> pick -> getNextState -> getGameEnded [-> getCanonicalForm]. */
typedef struct {
    int8_t*   planes;        /* in/out lane tiles */
    int       n_lanes;
    int       plies;         /* plies to play per lane in this call */
    uint64_t  seed;
    uint32_t  game_base;
    uint32_t* episodes;      /* [L] in/out (NULL: every lane starts at episode 0 and the count is dropped) */
    uint8_t*  players;       /* [L] in/out player to move; required when rotate=0, may be NULL when rotate=1 */
    int       rotate;        /* 1: keep the state canonical (player to move is always index 0) */
    int32_t*  first_plies;   /* [L] plies of the first game each lane finished in this call (0: none); may be NULL */
    float*    first_result;  /* [L][n] its check_end_game vector (in the stored frame); may be NULL */
    unsigned long long* counters; /* [2] += finished games, += plies played (may be NULL) */
} spl_rollout_args;

int spl_rollout(spl_ctx* ctx, const spl_rollout_args* args, void* stream);

/* get_score (:217-220) and get_round (:397-398) for every lane: scores int32[L][n], rounds int32[L] */
int spl_scores(spl_ctx* ctx, const int8_t* planes, int n_lanes, int32_t* scores, int32_t* rounds, void* stream);

/* get_symmetries (:349-395) on AoS input: for each of L states emits up to 1+9+2n variants.
 * out_states int8[L][18][R*7], out_pi float[L][18][406], out_valids uint8[L][18][406], out_count int32[L] */
int spl_symmetries(spl_ctx* ctx, const int8_t* aos, const float* pi, const uint8_t* valids, int n_lanes,
                   int8_t* out_states, float* out_pi, uint8_t* out_valids, int32_t* out_count, void* stream);

/* ======================================================================================================
 * MCTS tree arena  (replaces MCTS.py:16-250)
 *
 * One tree per game lane, one warp per tree. The reference's `nodes_data` dictionary (exact state bytes -> node; a DAG
 * with transpositions that persists across moves until reset_all_search_trees, MCTS.py:36,119-120,188-192) is a per-tree
 * hash table over node records that all trees allocate from one shared pool of 32 KB pages, inside one caller-owned
 * device buffer. A move is searched in waves:
 *
 *     spl_mcts_begin                      root lookup / creation, tree cleaning            (getActionProb :45-58)
 *     repeat: spl_mcts_select             descend by PUCT to an unevaluated node           (search :99-166, :199-237)
 *             <network on leaf rows>      caller-side (torch / any device code)            (nnet.predict :138)
 *             spl_mcts_expand             store Ps, root noise, back the value up          (:141-148, :168-177)
 *     spl_mcts_policy                     counts -> pruning -> temperature -> probs, q     (:61-97)
 *
 * Every tree runs exactly the reference's sequential algorithm (one leaf per tree per wave); the batch is the number
 * of trees. Simulations that end in a terminal node are backed up inside spl_mcts_select without a network call.
 * ====================================================================================================== */
typedef struct spl_mcts spl_mcts;

#define SPL_MCTS_MOVE_FORCED 1u   /* forced playouts + policy-target pruning for this move (:56, :69-74) */
#define SPL_MCTS_MOVE_NOISE 2u    /* root softmax + Dirichlet noise on the first simulation (:58, :141-143, :150-154) */

/* tree status bits (low byte of info[5] of spl_mcts_root_stats). A tree that reaches its node limit, or a shared pool
 * that runs dry, mid-move stops that tree's search early (the policy then reflects the simulations done so far); the
 * next spl_mcts_begin makes room, clears the bit and counts the event as a truncated search. */
#define SPL_MCTS_ST_OVERFLOW_NODES 1  /* the tree holds node_limit nodes */
#define SPL_MCTS_ST_OVERFLOW_POOL 2   /* no free page left in the shared pool */
#define SPL_MCTS_ST_PROTOCOL 4        /* select called while a leaf was still waiting for spl_mcts_expand */
#define SPL_MCTS_ST_BAD_STATE 8       /* a board the rules cannot have produced (a row that is no card of the tables, a deck count that is not
                                         its mask's popcount): node records hold a compact form that cannot represent it */

typedef struct {
    double   cpuct, fpu;        /* args.cpuct, args.fpu (pick_highest_UCB :199-213) */
    double   temperature0;      /* args.temperature[0]: root softmax before the noise (:141, :244-250) */
    double   dirichlet_alpha;   /* args.dirichletAlpha (:181); used by the on-device sampler when dir_values == NULL */
    uint64_t seed;              /* Philox key of the on-device Dirichlet sampler, counter (game, episode, root ply, 16 + 128 * action rank + draw) */
    uint32_t game_base;         /* game id of tree 0 */
    int      edge_reserve;      /* unused (kept for layout compatibility) */
    int      gc_reachable;      /* 0: exact cleaning (default) - drops only nodes no later search of the game can look up again:
                                   below the root's ply (like the reference's own cleaning :80-85) or holding a deck the root's
                                   deck is not a subset of (moves inside the tree never reveal cards, :228); result-neutral.
                                   1: a tree at its node limit keeps only what the new root reaches (not result-neutral) */
    int      rounds;            /* (descend, rules, attach) passes per selection wave (default 1): a descent that runs into
                                   a transposition or a terminal node continues in the next pass / wave; result-neutral */
    int      max_levels;        /* edges one descend call walks before it yields to the next wave (0: no limit); bounds the
                                   wave's latency by the typical, not the deepest, path; result-neutral */
} spl_mcts_params;

/* bytes of one node record with n_edges legal actions (32-byte header + compact state of 80 / 96 / 128 bytes + 24 bytes per edge,
 * rounded to 32): for sizing the pool */
size_t spl_mcts_record_bytes(int n_players, int n_edges);
/* bytes of device memory an arena needs. node_limit: most nodes ONE tree may hold (sizes its hash table); pool_bytes: the page
 * pool all trees share - size it for the AVERAGE tree (a tree is retired whenever a real move reveals a card) */
size_t spl_mcts_arena_bytes(int n_players, int n_trees, int node_limit, size_t pool_bytes, int leaves_per_tree);
/* MCTS.__init__ (:21-43): `arena` is caller-owned device memory of at least spl_mcts_arena_bytes, 256-byte aligned.
 * Call spl_mcts_reset(m, NULL, stream) once before the first spl_mcts_begin (it fills the ring of free pages). */
int  spl_mcts_create(spl_ctx* ctx, int n_trees, int node_limit, size_t pool_bytes, int leaves_per_tree, void* arena, size_t arena_bytes, spl_mcts** out);
/* leaves_per_tree (1..4): simulations one tree may have in flight per wave. 1 = the reference's sequential search (parity mode: visit
 * counts identical to MCTS.py). > 1 = virtual-loss leaf batching, NOT the reference's algorithm: every simulation in flight counts as
 * a lost visit (value -1) on the edges it walked until its value is backed up, so the simulations of one wave spread over different
 * leaves; the per-tree row buffers (leaf_states, leaf_valids, leaf_flags, pi, v) then hold n_trees * leaves_per_tree rows, row =
 * tree * leaves_per_tree + slot. One warp owns a tree, so the updates of Ns / Nsa / Qsa need no atomics. */
/* the lanes' episode counters (device uint32[T], read by spl_mcts_begin; may be NULL = 0): part of the key of the on-device
 * Dirichlet sampler, so that every episode of a lane draws its own root noise */
int  spl_mcts_set_episodes(spl_mcts* m, const uint32_t* episodes);
/* out4 (HOST int32[4]): pages in the pool, free now, fewest free since the last full reset, bytes per page (synchronises) */
int  spl_mcts_pool_stats(spl_mcts* m, int32_t* out4, void* stream);
void spl_mcts_destroy(spl_mcts* m);
int  spl_mcts_set_params(spl_mcts* m, const spl_mcts_params* p);
/* reset_all_search_trees (:188-192); tree_select (may be NULL) restricts it to trees with a non-zero byte */
int  spl_mcts_reset(spl_mcts* m, const uint8_t* tree_select, void* stream);
/* cleaning between waves, in any state of the search: every tree that holds more than fill_percent of node_limit copies what
 * it can still use (per gc_reachable) into fresh pages, gives the old ones back and re-bases the simulation in flight. */
int  spl_mcts_clean(spl_mcts* m, int fill_percent, void* stream);
/* start of getActionProb for every (selected) tree: roots int8[T][R*7] canonical boards, sims int32[T] simulation budget
 * (numMCTSSims or numMCTSSims // ratio_fullMCTS, :55), move_flags uint8[T] of SPL_MCTS_MOVE_* */
int  spl_mcts_begin(spl_mcts* m, const int8_t* roots, const int32_t* sims, const uint8_t* move_flags, const uint8_t* tree_select,
                    const double* dir_values, void* stream);
/* dir_values of spl_mcts_begin / spl_mcts_expand (may be NULL): double[T][406], the vector rng.dirichlet would return for
 * the tree's root, one value per legal action in action order (parity runs); NULL = the on-device Philox sampler.
 *
 * one selection wave = `rounds` x (descend, rules, attach) kernels. Outputs per tree: leaf_states int8[T][R*7],
 * leaf_valids uint8[T][406] (getValidMoves of the leaf), leaf_flags uint8[T] (1: this row needs the network).
 * A tree finishes at most one simulation per wave and may need an extra wave when a descent crosses several
 * transpositions / terminal nodes: run waves until counters[1] stays 0.
 * counters (may be NULL): int32[2], [0] += rows that need the network, [1] += trees whose budget is not yet spent. */
int  spl_mcts_select(spl_mcts* m, int8_t* leaf_states, uint8_t* leaf_valids, uint8_t* leaf_flags, int32_t* counters, void* stream);
/* pi float[T][406] = the network's probabilities (exp of the masked log-softmax, GenericNNetWrapper.py:166), v float[T][n] */
int  spl_mcts_expand(spl_mcts* m, const float* pi, const float* v, const double* dir_values, void* stream);
/* spl_mcts_expand followed by spl_mcts_select with the expansion and the next descent fused into one launch (the steady state
 * of a search: network -> expand_select -> network -> ...) */
int  spl_mcts_expand_select(spl_mcts* m, const float* pi, const float* v, const double* dir_values, int8_t* leaf_states,
                            uint8_t* leaf_valids, uint8_t* leaf_flags, int32_t* counters, void* stream);
/* The steady-state wave with the fused evaluator (spl_nnet_*) inside, three launches (four beyond 6144 trees, where the rules
 * step is a kernel of its own):
 *     expansion of the previous wave's leaves from pi / v + next descent -> rules -> { attach  ||  network -> pi, v }
 * The network starts as soon as the child states exist (it reads them where the rules kernel left them) and runs on an
 * internal side stream next to the attach kernel; the call joins it back into `stream` (CUDA-graph capturable). pi float[T][406]
 * and v float[T][n] are in/out: on entry the network outputs of the leaves selected by the previous wave (spl_mcts_select /
 * spl_mcts_expand_select + spl_nnet_forward, or the previous spl_mcts_wave_nnet), on return those of the new leaves.
 * nnet_blob = device copy of the packed weights (spl_nnet_pack). Same results as expand_select + spl_nnet_forward. */
int  spl_mcts_wave_nnet(spl_mcts* m, const void* nnet_blob, float* pi, float* v, const double* dir_values, int8_t* leaf_states,
                        uint8_t* leaf_valids, uint8_t* leaf_flags, int32_t* counters, void* stream);
/* getActionProb's tail: probs double[T][406], q double[T][n]; temp == 0 gives the one-hot of the FIRST most visited action */
int  spl_mcts_policy(spl_mcts* m, double temp, double* probs, double* q, void* stream);
/* Self-play with lanes that move on their own: for every tree whose budget is spent (simulations done >= budget, or a status bit
 * set), getActionProb's tail and the caller's draw from it (Coach.py:75-86) in one launch, without materialising the 406
 * probabilities: actions int16[T] = the action drawn with probability proportional to the (pruned, tempered) visit counts, -1 for
 * a tree that is still searching; finished uint8[T]; counters (may be NULL) int64[2]: [0] += simulations of the finished trees,
 * [1] += finished trees. The uniform comes from Philox keyed (seed, game, episode, root ply) like the environment's chance events:
 * episodes uint32[T] = the lanes' episode counters (NULL: 0). temp == 0: the first most visited action. */
int  spl_mcts_sample_moves(spl_mcts* m, double temp, const uint32_t* episodes, int16_t* actions, uint8_t* finished, long long* counters,
                            void* stream);
/* raw root statistics, any pointer may be NULL: nsa int32[T][406], qsa double[T][406] (-42 = unvisited), ps float[T][406],
 * info int32[T][16] = nodes, edges, root Ns, simulations done, network calls since reset, status bits | truncated searches << 8,
 *                     lossy resets * 65536 + cleanings, root Qs (float bits), then 4 floats (bits): the value vector the last
 *                     finished simulation returned at the root (what MCTS.search returns, :99-177), then the sum of the
 *                     path lengths of the simulations since the last reset; 3 spare words */
int  spl_mcts_root_stats(spl_mcts* m, int32_t* nsa, double* qsa, float* ps, int32_t* info, void* stream);
/* diagnostics: per-tree time stamps of the wave kernels into stamps int64[T][16] (device memory; NULL switches it off again):
 * expand+descend kernel [0] globaltimer at start, [1..3] SM clock at start / after the expansion / after the descent, [4] result code
 * * 1000 + path length, [5] globaltimer at the end; rules kernel (first tree of each warp) [6] globaltimer, [7..10] SM clock at
 * start / states loaded / rules done / end, [11] globaltimer; attach kernel [12] globaltimer, [13..14] SM clock, [15] globaltimer */
int  spl_mcts_debug_profile(spl_mcts* m, long long* stamps);
/* deterministic stand-in network ("fixed NN outputs"): a pure function of the state bytes with exact dyadic outputs; the
 * golden MCTS fixtures were produced by the reference's own MCTS.py with this function as its network */
int  spl_mcts_fixed_net(spl_ctx* ctx, const int8_t* states, const uint8_t* valids, int n_rows, float* pi, float* v, void* stream);

/* ======================================================================================================
 * Fused leaf evaluator  (replaces GenericNNetWrapper.predict :141-168 + SplendorNNet.forward, SplendorNNet.py:127-159,
 * for inference on device rows; one launch per batch, bf16 tensor-core products with fp32 accumulation)
 * ====================================================================================================== */
/* bytes of the packed weight blob for an n-player network */
size_t spl_nnet_blob_bytes(int n_players);
/* HOST-side packing (no GPU involved): folds the BatchNorm layers (eval mode) and lays the weights out as the kernel
 * streams them. `tensors`: 46 host float arrays in the reference's state_dict order (SplendorNNet.py:75-114), BatchNorm
 * entries as (weight, bias, running_mean, running_var):
 *   dense2d_1.0.{weight,bias}, dense2d_1.1.{4}, dense2d_1.3.{w,b}, partialgpool_1.dense_part.0.{w,b}, partialgpool_1.dense_part.1.{4},
 *   dense2d_3.0.{w,b}, dense1d_4.0.{w,b}, partialgpool_4.dense_part.0.{w,b}, partialgpool_4.dense_part.1.{4}, dense1d_5.0.{w,b},
 *   dense1d_5.1.{4}, dense1d_5.3.{w,b}, partialgpool_5.dense_part.0.{w,b}, partialgpool_5.dense_part.1.{4},
 *   output_layers_PI.0.{w,b}, output_layers_PI.1.{w,b}, output_layers_V.0.{w,b}, output_layers_V.1.{w,b} */
int spl_nnet_pack(int n_players, const float* const* tensors, void* blob_host, size_t blob_bytes);
/* states int8[B][R*7], valids uint8[B][406] -> pi float[B][406] (probabilities), v float[B][n]; blob = device copy of the
 * packed weights (16-byte aligned) */
int spl_nnet_forward(spl_ctx* ctx, const void* blob, const int8_t* states, const uint8_t* valids, int n_rows, float* pi, float* v,
                     void* stream);
/* self-test of the tcgen05 / TMEM building blocks the evaluator is made of: out float[128][n] = A bf16[128][k] . B bf16[n][k]^T on
 * one CTA (n a multiple of 32, k a multiple of 16, both <= 256; device pointers). *err_flag (device int) becomes 1 if the completion barrier timed out. */
int spl_umma_selftest(spl_ctx* ctx, const void* a_bf16, const void* b_bf16, float* out, int n, int k, int* err_flag, void* stream);
/* the same product with B staged MN-major in shared memory (element (n, k) at (n/8) stride_n8 + (k/8) stride_k8 + (k%8) 16 + (n%8) 2 bytes;
   lbo / sbo = the two byte offsets of the shared-memory descriptor): the operand layout of the transposed evaluator (spl_nnet.cu) */
int spl_umma_selftest_mn(spl_ctx* ctx, const void* a_bf16, const void* b_bf16, float* out, int n, int k, int stride_k8, int stride_n8, int lbo,
                         int sbo, int* err_flag, void* stream);
/* diagnostics: SM cycles per CTA to stream `tiles` tiles of tile_bytes from an L2-resident buffer of src_tiles tiles into a ring of `depth`
   shared-memory slots, all `grid` CTAs at once (mode 0: cp.async.bulk by one thread, 1: 16-byte cp.async by one warp); out[grid] on the device */
int spl_umma_stream_cycles(spl_ctx* ctx, const void* src, int src_tiles, int tile_bytes, int depth, int tiles, int mode, int grid, long long* out,
                           void* stream);
/* diagnostics: SM cycles of reps x ksteps tcgen05.mma (M 128, N n, K 16) on zero operands: out2[0] issue, out2[1] until the commit arrives
   (device memory); b_mn = 1: B operand MN-major with the given SBO */
int spl_umma_mma_cycles(spl_ctx* ctx, int n, int ksteps, int b_mn, int sbo, int reps, long long* out2, void* stream);
/* diagnostics: SM-clock time stamps of the phases of CTA 0 in the last spl_nnet_forward launch (long long[32]) */
int spl_nnet_debug_stamps(long long* out32);
/* diagnostics: per weight tile of CTA 0 in the last launch: SM clock when it was requested, when it had landed, when its MMAs were issued (long long[3][48]) */
int spl_nnet_debug_tile_stamps(long long* out144);
/* diagnostics: globaltimer (ns) at the start and the end of the first 160 CTAs of the last spl_nnet_forward launch (long long[320]) */
int spl_nnet_debug_cta_times(long long* out320);

#ifdef __cplusplus
}
#endif
#endif
