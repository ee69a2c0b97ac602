// TEST-ONLY: compiles the device rules core (csrc/spl_rules.cuh) with g++ so its logic can be
// checked against the oracle in the GPU-less build container. Never loaded by the product.
#include <stdint.h>
#include <string.h>
#include "../../alphazero-general-ori_b200/csrc/spl_rules.cuh"

struct Aos {
    int8_t* p;
    int get(int row, int col) const { return p[7 * row + col]; }
    void set(int row, int col, int v) { p[7 * row + col] = (int8_t)v; }
};

#define DISPATCH(n, CALL) switch (n) { case 2: { constexpr int N = 2; CALL; } break; case 3: { constexpr int N = 3; CALL; } break; default: { constexpr int N = 4; CALL; } }

extern "C" {
void hs_valid_mask(int n, const int8_t* st, int player, int limit, uint32_t flags, uint32_t* m13) {
    Aos s{(int8_t*)st}; SplRules r{limit, flags};
    DISPATCH(n, spl_valid_mask<N>(s, player, r, m13));
}
int hs_apply_move(int n, int8_t* st, int a, int player, int mode, int code, uint64_t seed, uint32_t game, uint32_t episode) {
    Aos s{st}; SplChance ch{mode, code, seed, game, episode, (uint32_t)(uint8_t)st[6]};
    int rc = 0;
    DISPATCH(n, rc = spl_apply_move<N>(s, a, player, ch));
    return rc;
}
int hs_game_ended(int n, const int8_t* st, int limit, uint32_t flags, float* out) {
    Aos s{(int8_t*)st}; SplRules r{limit, flags}; bool e = false;
    DISPATCH(n, e = spl_game_ended<N>(s, r, out));
    return e;
}
void hs_rotate(int n, int8_t* st, int k, uint32_t flags) {
    Aos s{st}; SplRules r{10, flags};
    DISPATCH(n, spl_rotate<N>(s, k, r));
}
void hs_init_explicit(int n, int8_t* st, const uint8_t* deals, const uint8_t* nobles) {
    Aos s{st};
    DISPATCH(n, spl_init_explicit<N>(s, deals, nobles));
}
void hs_init_philox(int n, int8_t* st, uint64_t seed, uint32_t game, uint32_t episode) {
    Aos s{st};
    DISPATCH(n, spl_init_philox<N>(s, seed, game, episode));
}
int hs_pick_random(const uint32_t* m13, uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply) {
    return spl_pick_random(m13, seed, game, episode, ply);
}
int hs_score(int n, const int8_t* st, int p, uint32_t flags) {
    Aos s{(int8_t*)st}; SplRules r{10, flags}; int v = 0;
    DISPATCH(n, v = spl_score<N>(s, p, r));
    return v;
}
}
