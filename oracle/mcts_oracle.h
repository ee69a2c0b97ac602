/* TEST INFRASTRUCTURE - CPU oracle for the search half of the hot path: a plain-C restatement of the
 * reference's MCTS.py (class MCTS :16-192, pick_highest_UCB :199-219,
 * get_next_best_action_and_canonical_state :222-237, np_roll :194, normalise :239, softmax :244) on top of the
 * rules oracle (splendor_oracle.h). Only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
 *
 * Pinning: tests/golden/mcts_*.npz hold outputs of the reference's own MCTS.py run in the build container
 * (oracle/refgen/gen_mcts_golden.py) with the fixed network of oracle/fakenn.py; tests/test_mcts_oracle.py
 * compares visit counts (exact), Qsa (1e-12), probs and q against them.
 */
#ifndef MCTS_ORACLE_H
#define MCTS_ORACLE_H
#include <stdint.h>
#include "splendor_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int    num_sims;          /* args.numMCTSSims */
    int    ratio_full;        /* args.ratio_fullMCTS */
    int    forced_playouts;   /* args.forced_playouts */
    int    dirichlet_noise;   /* MCTS(..., dirichlet_noise=) */
    double cpuct, fpu;        /* args.cpuct, args.fpu */
    double temperature0;      /* args.temperature[0] (root softmax before the noise) */
} mo_args;

/* network callback: predict(board, valids) -> (Ps float32[406], v float32[n]) ; NULL = built-in oracle/fakenn.py */
typedef void (*mo_predict_fn)(const int8_t* state, const uint8_t* valids406, float* ps406, float* v_n, void* user);

typedef struct mo_tree mo_tree;

mo_tree* mo_create(const spo_rules* rules, const mo_args* args, mo_predict_fn fn, void* user);
void     mo_destroy(mo_tree* t);
void     mo_reset(mo_tree* t);                       /* reset_all_search_trees :188-192 */
long     mo_num_nodes(const mo_tree* t);
long     mo_nn_calls(const mo_tree* t);

/* getActionProb :45-97. full_search: the outcome of `force_full_search or rng.random() < prob_fullMCTS`;
 * dir_values: the vector rng.dirichlet would return (used at step 0 of a full search when dirichlet_noise), or NULL
 * to draw nothing (then noise must be off). temp==0 returns the one-hot of the FIRST best action.
 * outputs may be NULL. */
int mo_get_action_prob(mo_tree* t, const int8_t* canonical, double temp, int full_search, const double* dir_values,
                       double* probs406, double* q_n, int64_t* nsa406, double* qsa406, long* ns, float* qs);

/* search :99-177 - one simulation from `canonical`; v_out float32[n] */
void mo_search(mo_tree* t, const int8_t* canonical, int dirichlet_noise, int forced_playouts, int step,
               const double* dir_values, float* v_out);

/* the fixed network of oracle/fakenn.py, also usable as a callback */
void mo_fake_predict(const int8_t* state, int state_bytes, const uint8_t* valids406, int n_players, float* ps406, float* v_n);

#ifdef __cplusplus
}
#endif
#endif
