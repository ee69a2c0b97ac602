/* TEST INFRASTRUCTURE - CPU oracle for the Splendor rules hot path.
 *
 * A plain-C restatement of the reference's SplendorLogicNumba.py `Board` (patched per
 * SURVEY.md §8(c), see oracle/refgen/build_patched_ref.py). It is the checker for the CUDA
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it. The product never links or calls it.
 *
 * Pinning: the reference ships no golden vectors for this path. The oracle is
 * pinned against outputs of the reference itself run in the build container, committed under
 * tests/golden/ by oracle/refgen/gen_golden.py, and live against the patched reference when
 * /root/reference is present (tests/test_oracle_vs_reference_live.py).
 *
 * State = int8[R][7], R = 32 + 10n + n*n, exactly the reference's array (SplendorLogicNumba.py:291-303).
 */
#ifndef SPLENDOR_ORACLE_H
#define SPLENDOR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPO_ACTIONS 406

/* rule switches: Board fields ENABLE_ACTION_RESERVE / ENABLE_ACTION_GIVEBACK / NUM_TOKEN_LIMIT
 * (SplendorLogicNumba.py:96-98). ref_compat=1 reproduces the n>=3 quirks:
 * noble stride 3 in get_score/swap_players and int8(999) in judge(). */
typedef struct {
    int n_players;
    int token_limit;
    int enable_reserve;
    int enable_giveback;
    int ref_compat;
} spo_rules;

int  spo_rows(int n);                       /* observation_size :25-27 */
void spo_default_rules(spo_rules* r, int n);

/* init_game :222-246 without the chance part: bank, deck counters/bitmasks, everything else 0 */
void spo_init_empty(int8_t* st, const spo_rules* r);
/* explicit deck draw (= _get_select_card :423-443 followed by the slot write of _fill_new_card :445-450).
 * slot = 0..11 visible slot. returns 0, or -1 if that card is not in the deck */
int  spo_deal_to_slot(int8_t* st, const spo_rules* r, int slot, int color, int idx);
void spo_set_noble(int8_t* st, const spo_rules* r, int slot, int noble_id);  /* :241-243 */

void spo_valid_moves(const int8_t* st, const spo_rules* r, int player, uint8_t* out406); /* :251-265 */

/* make_move :267-289. reveal: -1 = deterministic (no draw), else color*8+idx = the card the
 * reference's RNG would have drawn (replay), or -2 = draw with the oracle's Philox stream
 * (seed, game, episode; ply is read from the state). Returns next player, or <0 on an
 * action the reference leaves undefined (no free reserve slot / empty deck draw). */
int  spo_make_move(int8_t* st, const spo_rules* r, int move, int player, int reveal,
                   uint64_t seed, uint32_t game, uint32_t episode);

void spo_check_end_game(const int8_t* st, const spo_rules* r, float* out_n); /* :320-334 */
int  spo_get_score(const int8_t* st, const spo_rules* r, int player);         /* :217-220 */
int  spo_get_round(const int8_t* st);                                         /* :397-398 */
void spo_swap_players(int8_t* st, const spo_rules* r, int nb_swaps);         /* :338-347 */

/* get_symmetries :349-395. outputs up to 1+9+2n <= 18 (state, policy, valids) triples; returns count */
int  spo_symmetries(const int8_t* st, const spo_rules* r, const float* pi, const uint8_t* valids,
                    int8_t* out_states, float* out_pi, uint8_t* out_valids);

/* ---- chance by counter-based Philox4x32-10 (the product's definition; restated here so the
 * CUDA sampler can be checked bit for bit). key=(seed lo, seed hi), counter=(game, episode, ply, stream) */
void spo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* which card a Philox draw from deck `tier` yields for this state: returns color*8+idx or -1 if empty */
int  spo_philox_draw(const int8_t* st, int tier, uint64_t seed, uint32_t game, uint32_t episode,
                     uint32_t ply, uint32_t stream);
/* full Philox game start: 12 deals (stream 2, ply field = slot) + nobles (stream 3) */
void spo_init_philox(int8_t* st, const spo_rules* r, uint64_t seed, uint32_t game, uint32_t episode);
/* uniform choice among set flags (stream 1): the rollout policy of the bench */
int  spo_philox_pick(const uint8_t* valids406, uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply);

/* random rollouts for the CPU baseline: plays `games` full games starting at game id `game0`,
 * returns total plies; optional per-game outputs */
long spo_rollout(const spo_rules* r, uint64_t seed, uint32_t game0, int games, int32_t* plies_out,
                 float* result_out);

#ifdef __cplusplus
}
#endif
#endif
