// Splendor rules core for the sm_100a kernels (device code; also compiled by g++ for the
// test-only host simulator in tests/hostsim/, never by the product's host path).
//
// What it replaces in the reference (SplendorLogicNumba.py, jitclass Board):
//   spl_valid_mask   <- valid_moves :251-265 and the nine _valid_* predicates :476-680
//   spl_apply_move   <- make_move :267-289 and _buy/_reserve/_buy_reserve/_get_gems/_give_gems/
//                       _give_and_get_gems/_reserve_and_give/_give_nobles_if_earned/_fill_new_card
//   spl_game_ended   <- check_end_game :320-334, get_score :217, judge :306-318
//   spl_rotate       <- swap_players :338-347
//   spl_init_*       <- init_game :222-246
//   spl_draw_philox  <- _get_deck_card :400-412 with the two uniforms taken from Philox4x32-10
//
// Everything is templated on the number of players N (row offsets become immediates) and on a
// state accessor S { int get(row, col) const; void set(row, col, int v); } so the same code
// runs on shared-memory byte planes (step kernels, one thread per game lane) and on
// array-of-struct node states (tree kernels).
//
// The mask is built bit-parallel: 30 "bank can supply" flags and 20+40 "player can give" flags are
// evaluated once with nibble-SWAR compares, the 345 exchange actions are table-ANDs of those.
#pragma once
#include <stdint.h>
#include <utility>

#ifdef __CUDACC__
#define SPL_D __device__ __forceinline__
#define SPL_COLD static __device__ __noinline__   // big, rarely or divergently executed blocks: kept out of the straight-line hot code (instruction cache)
#define SPL_CTABLE static __constant__
#define SPL_GTABLE static __device__ const
#define SPL_KTABLE static constexpr
#define SPL_FFS(x) __ffs(x)
#define SPL_POPC(x) __popc(x)
#define SPL_MULHI(a, b) __umulhi((a), (b))
#else
#define SPL_D static inline
#define SPL_COLD static
#define SPL_CTABLE static const
#define SPL_GTABLE static const
#define SPL_KTABLE static constexpr
#define SPL_FFS(x) __builtin_ffs(x)
#define SPL_POPC(x) __builtin_popcount(x)
#define SPL_MULHI(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#endif
#include "spl_tables.cuh"

#define SPL_ACTIONS 406
#define SPL_MASK_WORDS 13

// rule switches (Board fields ENABLE_ACTION_RESERVE / ENABLE_ACTION_GIVEBACK / NUM_TOKEN_LIMIT :96-98)
#define SPL_F_RESERVE 1u
#define SPL_F_GIVEBACK 2u
#define SPL_F_REFCOMPAT 4u   // reproduce the n>=3 quirks: noble stride 3 (:219,:345), int8(999) (:313)

struct SplRules {
    int limit;
    uint32_t flags;
};

template <int N>
struct SplLay {   // copy_state :291-303
    static constexpr int ROWS = 32 + 10 * N + N * N;
    static constexpr int CELLS = 7 * ROWS;
    static constexpr int BANK = 0, CARDS = 1, DECK = 25, NOBLES = 31;
    static constexpr int PGEMS = 32 + N, PNOBLES = 32 + 2 * N, PCARDS = 32 + 3 * N + N * N, PRES = 32 + 4 * N + N * N;
    static constexpr int NUM_NOBLES = N + 1;
    static constexpr int GEMS0 = (N == 2) ? 4 : (N == 3) ? 5 : 7;   // num_gems_in_play :90
    static constexpr int MAX_MOVES = 62 * N;                        // :92 (uint8)
};

enum { SPL_GOLD = 5, SPL_PTS = 6 };

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11). key = seed, counter = (game, episode, ply, stream).
// streams: 0 reveal during a move, 1 rollout-policy pick, 2 initial deal (ply = slot), 3 nobles
// ------------------------------------------------------------------------------------------
struct SplPhilox {
    uint32_t v[4];
};
SPL_COLD SplPhilox spl_philox(uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply, uint32_t stream) {
    uint32_t c0 = game, c1 = episode, c2 = ply, c3 = stream, k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t h0 = SPL_MULHI(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = SPL_MULHI(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    SplPhilox r;
    r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
    return r;
}

// chance source of one move. mode 0: deterministic (no reveal; the in-tree MCTS step), 1: replay
// the given outcome (colour*8+idx; parity runs feed the reference's own draws), 2: Philox
struct SplChance {
    int mode;
    int code;
    uint64_t seed;
    uint32_t game, episode, ply;
};

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
template <class S>
SPL_D int spl_sum5(const S& s, int row) {
    return s.get(row, 0) + s.get(row, 1) + s.get(row, 2) + s.get(row, 3) + s.get(row, 4);
}

// "the five costs sum to non-zero" (the reference's emptiness test): with non-negative costs that is "some cost is set";
// only if a negative cost shows up (never in a game) the exact sum is taken
template <class S>
SPL_D bool spl_sum5_nonzero(const S& s, int row, int or_of_costs) {
    if (or_of_costs >= 0) return or_of_costs != 0;
    return spl_sum5(s, row) != 0;
}

SPL_D uint32_t spl_nib5(const int* v) {   // clamp to [0,7] and pack one nibble per colour
    uint32_t x = 0;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        int t = v[c] < 0 ? 0 : (v[c] > 7 ? 7 : v[c]);
        x |= (uint32_t)t << (4 * c);
    }
    return x;
}
// nibble-wise have >= need for all five colours (have, need <= 7 per nibble)
SPL_D uint32_t spl_ge5(uint32_t have, uint32_t need) {
    return ((((have | 0x88888u) - need) & 0x88888u) == 0x88888u) ? 1u : 0u;
}

SPL_D void spl_mask_or(uint32_t* m, int pos, uint32_t bits, int nbits) {   // pos/nbits compile-time at every call site
    m[pos >> 5] |= bits << (pos & 31);
    if ((pos & 31) + nbits > 32) m[(pos >> 5) + 1] |= bits >> (32 - (pos & 31));
}

// exchanges 60..289 (_valid_exchange :634-669), word-parallel: action a is legal iff its take flag AND its give
// flag are set, so mask = OR_t(T_t ? EXM_TAKE[t]) & OR_g(G_g ? EXM_GIVE[g]) & regime. The tables are compile-time
// constants: after unrolling every (flag, word) pair with a non-zero table entry is one LOP3 with an immediate.
template <int F, int W>
SPL_D void spl_ex_take_fw(uint32_t* ts, uint32_t sel) {
    constexpr uint32_t v = SPL_EXM_TAKE[F][W];
    if constexpr (v != 0u) ts[W] |= sel & v;
}
template <int F, int W>
SPL_D void spl_ex_give_fw(uint32_t* gs, uint32_t sel) {
    constexpr uint32_t v = SPL_EXM_GIVE[F][W];
    if constexpr (v != 0u) gs[W] |= sel & v;
}
template <int F, int... W>
SPL_D void spl_ex_take_f(uint32_t* ts, uint32_t T, std::integer_sequence<int, W...>) {
    const uint32_t sel = 0u - ((T >> F) & 1u);
    (spl_ex_take_fw<F, W + 1>(ts, sel), ...);
}
template <int F, int... W>
SPL_D void spl_ex_give_f(uint32_t* gs, uint32_t G, std::integer_sequence<int, W...>) {
    const uint32_t sel = 0u - ((G >> F) & 1u);
    (spl_ex_give_fw<F, W + 1>(gs, sel), ...);
}
template <int... F>
SPL_D void spl_ex_take_all(uint32_t* ts, uint32_t T, std::integer_sequence<int, F...>) {
    (spl_ex_take_f<F>(ts, T, std::make_integer_sequence<int, 9>{}), ...);
}
template <int... F>
SPL_D void spl_ex_give_all(uint32_t* gs, uint32_t G, std::integer_sequence<int, F...>) {
    (spl_ex_give_f<F>(gs, G, std::make_integer_sequence<int, 9>{}), ...);
}
template <int W>
SPL_D uint32_t spl_ex_regime_w(int regime) {
    constexpr uint32_t r0 = SPL_EXM_REGIME[0][W], r1 = SPL_EXM_REGIME[1][W], r2 = SPL_EXM_REGIME[2][W];
    return regime == 0 ? r0 : (regime == 1 ? r1 : r2);
}
template <int... W>
SPL_D void spl_ex_combine(uint32_t* m, const uint32_t* ts, const uint32_t* gs, int regime, std::integer_sequence<int, W...>) {
    ((m[W + 1] |= ts[W + 1] & gs[W + 1] & spl_ex_regime_w<W + 1>(regime)), ...);
}
SPL_D void spl_ex_words(uint32_t* m, uint32_t T, uint32_t G, int regime) {   // mask words 1..9 hold actions 60..289
    uint32_t ts[SPL_MASK_WORDS], gs[SPL_MASK_WORDS];
#pragma unroll
    for (int w = 0; w < SPL_MASK_WORDS; w++) ts[w] = gs[w] = 0u;
    spl_ex_take_all(ts, T, std::make_integer_sequence<int, 30>{});
    spl_ex_give_all(gs, G, std::make_integer_sequence<int, 20>{});
    spl_ex_combine(m, ts, gs, regime, std::make_integer_sequence<int, 9>{});
}

// ------------------------------------------------------------------------------------------
// legality mask (406 bits, little-endian in 13 words) for `player`
// ------------------------------------------------------------------------------------------
// candidate card i of valid_moves (0..11 visible, 12..14 the player's reserved slots): bit i of `buyable` (_valid_buy :476-501,
// _valid_buy_reserve :538-552) and of `present` (:511). have[c] = gems + bonuses of colour c, gold = the player's gold.
template <int N, class S>
SPL_D void spl_card_bits(const S& s, int p, int i, const int* have, int gold, uint32_t& buyable, uint32_t& present) {
    typedef SplLay<N> L;
    const int row = i < 12 ? L::CARDS + 2 * i : L::PRES + 6 * p + 2 * (i - 12);
    int missing = 0, any = 0;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        const int cost = s.get(row, c);
        const int d = (int)(int8_t)(cost - have[c]);
        missing += d > 0 ? d : 0;
        any |= cost;
    }
    const bool card = spl_sum5_nonzero(s, row, any);
    buyable |= (uint32_t)((missing <= gold) && card) << i;
    present |= (uint32_t)card << i;
}

// the generated masks as immediates (the tables are constexpr: read them in constant expressions only)
template <int C>
SPL_D void spl_combo_fail(const int* b, const int* g, uint32_t& t_fail, uint32_t& g_fail) {
    constexpr uint32_t M = SPL_COMBO_WITH[C];
    t_fail |= b[C] >= 1 ? 0u : M;
    g_fail |= g[C] >= 1 ? 0u : M;
    if constexpr (C < 4) spl_combo_fail<C + 1>(b, g, t_fail, g_fail);
}
template <int C>
SPL_D void spl_give3_fail(const int* g, uint64_t& fail) {
    constexpr uint64_t M1 = SPL_GIVE3_NEEDS[C][0], M2 = SPL_GIVE3_NEEDS[C][1], M3 = SPL_GIVE3_NEEDS[C][2];
    fail |= (g[C] >= 1 ? 0ull : M1) | (g[C] >= 2 ? 0ull : M2) | (g[C] >= 3 ? 0ull : M3);
    if constexpr (C < 4) spl_give3_fail<C + 1>(g, fail);
}

// pre (may be NULL): {buyable, present} of the 15 candidate cards, computed elsewhere by spl_card_bits (the tree kernels spread
// them over the lanes of a warp)
template <int N, class S>
SPL_D void spl_valid_mask(const S& s, int p, SplRules r, uint32_t* m, const uint32_t* pre = nullptr) {
    typedef SplLay<N> L;
#pragma unroll
    for (int w = 0; w < SPL_MASK_WORDS; w++) m[w] = 0;

    int b[6], g[7], pc[5];
#pragma unroll
    for (int c = 0; c < 6; c++) b[c] = s.get(L::BANK, c);
#pragma unroll
    for (int c = 0; c < 7; c++) g[c] = s.get(L::PGEMS + p, c);
#pragma unroll
    for (int c = 0; c < 5; c++) pc[c] = s.get(L::PCARDS + p, c);
    const int tokens = g[0] + g[1] + g[2] + g[3] + g[4] + g[5] + g[6];   // players_gems[player].sum() incl. gold
    const int gold = g[5];

    // --- buy visible 0..11 (_valid_buy :476-501), card presence for reserve (:511), buy reserved 27..29 (_valid_buy_reserve
    // :538-552): one rolled loop over the 15 candidate cards (bit i of `buyable` / `present`)
    uint32_t buyable = 0, present = 0;
    int have[5];   // gems + bonuses per colour; the reference subtracts both from the cost in int8 arithmetic (wraps mod 256)
#pragma unroll
    for (int c = 0; c < 5; c++) have[c] = g[c] + pc[c];
    if (pre) {
        buyable = pre[0]; present = pre[1];
    } else {
#ifndef SPL_CARD_UNROLL
#define SPL_CARD_UNROLL 1
#endif
        constexpr int kCardUnroll = SPL_CARD_UNROLL;
#pragma unroll kCardUnroll
        for (int i = 0; i < 15; i++) spl_card_bits<N>(s, p, i, have, gold, buyable, present);
    }
    const uint32_t buy = buyable & 0xFFFu, buyres = (buyable >> 12) & 7u;
    present &= 0xFFFu;
    // --- reserve 12..26 (_valid_reserve :508-515)
#pragma unroll
    for (int t = 0; t < 3; t++) present |= (uint32_t)(spl_sum5(s, L::DECK + 2 * t) != 0) << (12 + t);
    const bool third_free = spl_sum5(s, L::PRES + 6 * p + 5) == 0;   // GAIN row of the third slot
    const uint32_t rsv_nolimit = third_free ? present : 0u;
    const bool rsv_blocked = !(r.flags & SPL_F_RESERVE) || (tokens == r.limit && b[SPL_GOLD] > 0);
    const uint32_t rsv = rsv_blocked ? 0u : rsv_nolimit;

    // --- bank-can-supply flags T (30) and player-can-give flags G (20)
    bool bank_neg = false;
#pragma unroll
    for (int c = 0; c < 5; c++) bank_neg |= b[c] < 0;
    // The 25 "different gems" rows need one gem of each colour of the row: a subset test on the 5-bit set of colours
    // the bank (or the player) holds. Rows 25..29 / 15..19 ("identical") look at one colour's count.
    // _valid_get_gems, is_limit=False :562-568; _valid_give_gems :595 (first 15 rows): a row is possible iff no colour it needs is missing,
    // rows = ~OR_{c missing} SPL_COMBO_WITH[c] (compile-time masks: five selects per side instead of 25 subset tests)
    uint32_t t_fail = 0, g_fail = 0;
    bool g_neg = false;
    spl_combo_fail<0>(b, g, t_fail, g_fail);
#pragma unroll
    for (int c = 0; c < 5; c++) g_neg |= g[c] < 0;
    uint32_t T = ~t_fail & 0x1FFFFFFu, G = ~g_fail & 0x7FFFu;
    if (bank_neg) T = 0;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        T |= (uint32_t)(b[c] >= 4) << (25 + c);     // _valid_get_gems_identical :578-583
        G |= (uint32_t)(g[c] >= 2) << (15 + c);     // _valid_give_gems_identical :609
    }
    if (r.flags & SPL_F_GIVEBACK) {
        if (g_neg) G &= 0xF8000u;   // a negative count fails every "different gems" row; identical rows look at one colour
    } else {
        G = 0;
    }
    // --- take only 30..59 (:256, limits :567-574)
    int nspec = 0;
#pragma unroll
    for (int c = 0; c < 5; c++) nspec += b[c] != 0;
    uint32_t take = 0;
    if (tokens + 1 <= r.limit && (tokens == 9 || nspec == 1)) take |= T & 0x1Fu;
    if (tokens + 2 <= r.limit && (tokens == 8 || nspec == 2)) take |= T & 0x7FE0u;
    if (tokens + 3 <= r.limit) take |= T & 0x1FF8000u;
    if (tokens + 2 <= r.limit) take |= T & 0x3E000000u;

    spl_mask_or(m, 0, buy, 12);
    spl_mask_or(m, 12, rsv, 15);
    spl_mask_or(m, 27, buyres, 3);
    spl_mask_or(m, 30, take, 30);

    // --- exchanges 60..404 (_valid_exchange :615-680): the regime is chosen by the token total
    if (tokens > 7) {
        const int regime = tokens == r.limit - 2 ? 0 : (tokens == r.limit - 1 ? 1 : 2);
        spl_ex_words(m, T, G, regime);
        if (regime == 2) {
            if ((r.flags & SPL_F_GIVEBACK) && !g_neg) {   // take 3 / give 3 (:672, _valid_give_gems3 :602-607): actions 365..404
                // pattern 4 tk + q is legal iff take row 15 + tk is possible and the player holds what the pattern gives: the patterns
                // that fail are the OR, over the (colour, count) pairs the player cannot supply, of SPL_GIVE3_NEEDS (compile-time masks)
                uint64_t g3_fail = 0, on = 0;
                spl_give3_fail<0>(g, g3_fail);
#pragma unroll
                for (int tk = 0; tk < 10; tk++) on |= (uint64_t)((T >> (15 + tk)) & 1u) << (4 * tk);
                const uint64_t g3 = (on * 15ull) & ~g3_fail;
                m[11] |= (uint32_t)(g3 << 13);    // 365 = 11 * 32 + 13
                m[12] |= (uint32_t)(g3 >> 19);
            }
            if (b[SPL_GOLD] > 0) {                    // reserve + give 1 (:674-678): actions 290 + 5 i + colour, 75 bits from bit 2 of word 9
                const uint32_t grp = G & 0x1Fu;
                uint32_t lo = 0, mid = 0, hi = 0;     // bits 0..31 | 32..63 | 64..74 of the 75-bit field
#pragma unroll 1
                for (int i = 0; i < 15; i++) {
                    if (!((rsv_nolimit >> i) & 1u)) continue;
                    const int pos = 5 * i;
                    const uint64_t v = (uint64_t)grp << (pos & 31);
                    if (pos < 32) { lo |= (uint32_t)v; mid |= (uint32_t)(v >> 32); }
                    else if (pos < 64) { mid |= (uint32_t)v; hi |= (uint32_t)(v >> 32); }
                    else hi |= (uint32_t)v;
                }
                m[9] |= lo << 2;                       // 290 = 9 * 32 + 2
                m[10] |= (lo >> 30) | (mid << 2);
                m[11] |= (mid >> 30) | (hi << 2);
            }
        }
    }
    // --- pass 405 only when nothing else is legal (:263)
    uint32_t any = 0;
#pragma unroll
    for (int w = 0; w < SPL_MASK_WORDS; w++) any |= m[w];
    if (!any) m[12] |= 1u << (405 & 31);
}

// ------------------------------------------------------------------------------------------
// chance: draw from deck `tier` (count row 2t, MSB-first bitmask row 2t+1; :400-412)
// ------------------------------------------------------------------------------------------
// the two-stage choice of _get_deck_card (:400-412) from two 32-bit uniform words: colour with probability proportional
// to its remaining count, then uniformly among that colour's remaining cards. Returns colour*8+idx or -1 (empty deck).
template <int N, class S>
SPL_D int spl_draw_from(const S& s, int tier, uint32_t w0, uint32_t w1) {
    typedef SplLay<N> L;
    int cnt[5], total = 0;
#pragma unroll
    for (int c = 0; c < 5; c++) { cnt[c] = s.get(L::DECK + 2 * tier, c); total += cnt[c]; }
    if (total == 0) return -1;
    int k = (int)SPL_MULHI(w0, (uint32_t)total);
    int color = 4, acc = 0, ccount = cnt[4];
    bool found = false;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        acc += cnt[c];
        if (!found && acc > k) { color = c; ccount = cnt[c]; found = true; }
    }
    uint32_t bits = (uint32_t)(uint8_t)s.get(L::DECK + 2 * tier + 1, color);
    int j = (int)SPL_MULHI(w1, (uint32_t)ccount);
    int idx = -1;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (bits & (128u >> i)) {
            if (j == 0 && idx < 0) idx = i;
            j--;
        }
    }
    return idx < 0 ? -1 : color * 8 + idx;
}
template <int N, class S>
SPL_COLD int spl_draw_philox(const S& s, int tier, uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply, uint32_t stream) {
    const SplPhilox w = spl_philox(seed, game, episode, ply, stream);
    return spl_draw_from<N>(s, tier, w.v[0], w.v[1]);
}

// remove card (colour, idx) from deck `tier`; returns the packed card or 0 if it was not there
template <int N, class S>
SPL_D uint32_t spl_deck_take(S& s, int tier, int code) {
    typedef SplLay<N> L;
    const int color = code >> 3, idx = code & 7;
    uint32_t bits = (uint32_t)(uint8_t)s.get(L::DECK + 2 * tier + 1, color);
    uint32_t mk = 128u >> idx;
    if (!(bits & mk)) return 0u;
    s.set(L::DECK + 2 * tier + 1, color, (int)(int8_t)(uint8_t)(bits & ~mk));
    s.set(L::DECK + 2 * tier, color, s.get(L::DECK + 2 * tier, color) - 1);
    return SPL_CARDS[tier][color][idx];
}

template <class S>
SPL_D void spl_write_card(S& s, int row, uint32_t pk) {   // two rows: cost, gain (one-hot colour + points)
    const int col = (int)((pk >> 20) & 7u), pts = (int)((pk >> 24) & 15u);
#pragma unroll
    for (int c = 0; c < 5; c++) {
        s.set(row, c, pk ? (int)((pk >> (4 * c)) & 15u) : 0);
        s.set(row + 1, c, (pk && col == c) ? 1 : 0);
    }
    s.set(row, 5, 0); s.set(row, 6, 0);
    s.set(row + 1, 5, 0); s.set(row + 1, 6, pk ? pts : 0);
}

// draws per the chance source; returns packed card (0 = nothing drawn)
template <int N, class S>
SPL_D uint32_t spl_draw(S& s, int tier, const SplChance& ch) {
    typedef SplLay<N> L;
    if (ch.mode == 0) return 0u;
    if (spl_sum5(s, L::DECK + 2 * tier) == 0) return 0u;   // no more cards :402
    int code = ch.mode == 1 ? ch.code : spl_draw_philox<N>(s, tier, ch.seed, ch.game, ch.episode, ch.ply, 0);
    if (code < 0) return 0u;
    return spl_deck_take<N>(s, tier, code);
}

// ------------------------------------------------------------------------------------------
// move application
// ------------------------------------------------------------------------------------------
template <int N, class S>
SPL_D void spl_give_nobles(S& s, int p) {   // _give_nobles_if_earned :763-768 (all earned nobles at once)
    typedef SplLay<N> L;
    int pc[5], most = 0;
#pragma unroll
    for (int c = 0; c < 5; c++) { pc[c] = s.get(L::PCARDS + p, c); most = pc[c] > most ? pc[c] : most; }
    if (most < 3) return;   // every noble asks for at least 3 cards of some colour (SplendorLogic.py:320-332)
#pragma unroll 1
    for (int i = 0; i < L::NUM_NOBLES; i++) {
        int nb[7], tot = 0;
        bool ok = true;
#pragma unroll
        for (int c = 0; c < 7; c++) nb[c] = s.get(L::NOBLES + i, c);
#pragma unroll
        for (int c = 0; c < 5; c++) { tot += nb[c]; ok &= pc[c] >= nb[c]; }
        if (ok && tot > 0) {
#pragma unroll
            for (int c = 0; c < 7; c++) {
                s.set(L::PNOBLES + L::NUM_NOBLES * p + i, c, nb[c]);   // writer stride n+1 :767
                s.set(L::NOBLES + i, c, 0);
            }
        }
    }
}

template <int N, class S>
SPL_D void spl_buy_card(S& s, int cost_row, int p) {   // _buy_card :458-474
    typedef SplLay<N> L;
    int missing = 0;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        int cost = s.get(cost_row, c), gem = s.get(L::PGEMS + p, c), bon = s.get(L::PCARDS + p, c);
        int d = (int)(int8_t)(cost - gem - bon);
        missing += d > 0 ? d : 0;
        int need = (int)(int8_t)(cost - bon);
        need = need < 0 ? 0 : need;
        int paid = need < gem ? need : gem;
        s.set(L::PGEMS + p, c, gem - paid);
        s.set(L::BANK, c, s.get(L::BANK, c) + paid);
    }
    s.set(L::PGEMS + p, SPL_GOLD, s.get(L::PGEMS + p, SPL_GOLD) - missing);
    s.set(L::BANK, SPL_GOLD, s.get(L::BANK, SPL_GOLD) + missing);
#pragma unroll
    for (int c = 0; c < 7; c++) s.set(L::PCARDS + p, c, s.get(L::PCARDS + p, c) + s.get(cost_row + 1, c));   // :472
    spl_give_nobles<N>(s, p);
}

template <int N, class S>
SPL_D void spl_move_gems(S& s, int p, uint32_t nib, int sign) {   // _get_gems :585 (sign=+1) / _give_gems :685 (sign=-1)
    typedef SplLay<N> L;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        int k = (int)((nib >> (4 * c)) & 15u) * sign;
        if (k != 0) {
            s.set(L::BANK, c, s.get(L::BANK, c) - k);
            s.set(L::PGEMS + p, c, s.get(L::PGEMS + p, c) + k);
        }
    }
}

template <int N, class S>
SPL_D int spl_reserve(S& s, int i, int p, const SplChance& ch) {   // _reserve :517-536
    typedef SplLay<N> L;
    int slot = -1;
#pragma unroll
    for (int k = 2; k >= 0; k--)
        if (spl_sum5(s, L::PRES + 6 * p + 2 * k) == 0) slot = L::PRES + 6 * p + 2 * k;   // first free COST row :521
    if (slot < 0) return -2;   // undefined in the reference (unbound local); refuse instead of corrupting
    if (i < 12) {
#pragma unroll
        for (int c = 0; c < 7; c++) {
            s.set(slot, c, s.get(L::CARDS + 2 * i, c));
            s.set(slot + 1, c, s.get(L::CARDS + 2 * i + 1, c));
        }
        spl_write_card(s, L::CARDS + 2 * i, spl_draw<N>(s, i >> 2, ch));   // _fill_new_card :445-450
    } else if (ch.mode != 0) {   // top of deck; deterministic stores nothing :529-532
        uint32_t pk = spl_draw<N>(s, i - 12, ch);
        if (pk) spl_write_card(s, slot, pk);
    }
    if (s.get(L::BANK, SPL_GOLD) > 0) {   // :534-536
        s.set(L::PGEMS + p, SPL_GOLD, s.get(L::PGEMS + p, SPL_GOLD) + 1);
        s.set(L::BANK, SPL_GOLD, s.get(L::BANK, SPL_GOLD) - 1);
    }
    return 0;
}

// returns next player, or <0 for a move the reference leaves undefined. Pass (405) only bumps the ply (patch P6).
// Decode first, then ONE call site each for the card purchase, the reservation and the gem transfer: the lanes of a warp
// play different kinds of moves every ply, so every inlined copy of those blocks was executed (and fetched) by every warp.
template <int N, class S>
SPL_D int spl_apply_move(S& s, int a, int p, const SplChance& ch) {
    typedef SplLay<N> L;
    int buy_row = -1, res_idx = -1, comp_from = -1;
    uint32_t take = 0u, give = 0u;   // gems taken from / given back to the bank, one nibble per colour
    if (a < 12) {                                   // _buy :503-506
        buy_row = L::CARDS + 2 * a;
    } else if (a < 27) {                            // _reserve :517-536
        res_idx = a - 12;
    } else if (a < 30) {                            // _buy_reserve :554-560
        comp_from = a - 27;
        buy_row = L::PRES + 6 * p + 2 * comp_from;
    } else if (a < 60) {                            // _get_gems :585-593
        take = SPL_TAKE30_G[a - 30];
    } else if (a < 290 || (a >= 365 && a < 405)) {  // _give_and_get_gems :697-756
        take = SPL_TAKE30_G[SPL_EX_TAKE_G[a]];
        give = SPL_GIVE20_G[SPL_EX_GIVE_G[a]];
        if (a >= 365) give += SPL_GIVE20_G[SPL_EX_GIVE2_G[a]];   // two give operations; at most 3 of a colour in total
    } else if (a < 365) {                           // _reserve_and_give :759-761
        res_idx = (a - 290) / 5;
        give = 1u << (4 * ((a - 290) % 5));
    }
    if (buy_row >= 0) {
        spl_buy_card<N>(s, buy_row, p);
        if (comp_from < 0) {
            spl_write_card(s, buy_row, spl_draw<N>(s, a >> 2, ch));   // _fill_new_card :445-450
        } else {                                                      // close the gap in the reserve
            const int base = L::PRES + 6 * p;
            for (int k = comp_from; k < 2; k++) {
#pragma unroll
                for (int c = 0; c < 7; c++) {
                    s.set(base + 2 * k, c, s.get(base + 2 * k + 2, c));
                    s.set(base + 2 * k + 1, c, s.get(base + 2 * k + 3, c));
                }
            }
#pragma unroll
            for (int c = 0; c < 7; c++) { s.set(base + 4, c, 0); s.set(base + 5, c, 0); }
        }
    }
    if (res_idx >= 0) {
        const int rc = spl_reserve<N>(s, res_idx, p, ch);
        if (rc < 0) return rc;
    }
    if (take | give) {                              // _get_gems :585 then _give_gems :685 (the colour sets are disjoint)
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const int d = (int)((take >> (4 * c)) & 15u) - (int)((give >> (4 * c)) & 15u);
            if (d != 0) {
                s.set(L::BANK, c, s.get(L::BANK, c) - d);
                s.set(L::PGEMS + p, c, s.get(L::PGEMS + p, c) + d);
            }
        }
    }
    s.set(L::BANK, SPL_PTS, s.get(L::BANK, SPL_PTS) + 1);   // ply counter :287
    return (p + 1) % N;
}

// ------------------------------------------------------------------------------------------
// end of game
// ------------------------------------------------------------------------------------------
template <int N, class S>
SPL_D int spl_score(const S& s, int p, SplRules r) {   // get_score :217-220
    typedef SplLay<N> L;
    const int stride = (r.flags & SPL_F_REFCOMPAT) ? 3 : L::NUM_NOBLES;
    int pts = s.get(L::PCARDS + p, SPL_PTS);
    for (int i = 0; i < stride; i++) pts += s.get(L::PNOBLES + stride * p + i, SPL_PTS);
    return pts;
}

// returns true if the game is over; out[N] = 0 / +1 / -1 / 0.01
template <int N, class S>
SPL_D bool spl_game_ended(const S& s, SplRules r, float* out) {
    typedef SplLay<N> L;
#pragma unroll
    for (int p = 0; p < N; p++) out[p] = 0.f;
    const int round = (int)(uint8_t)s.get(L::BANK, SPL_PTS);   // get_round :397 reads uint8
    if (round % N != 0) return false;                           // :322
    int sc[N], smax = -128;
#pragma unroll
    for (int p = 0; p < N; p++) { sc[p] = (int)(int8_t)spl_score<N>(s, p, r); smax = sc[p] > smax ? sc[p] : smax; }
    if (!(smax >= 15 || round >= (int)(uint8_t)L::MAX_MOVES)) return false;
    int nmax = 0;
#pragma unroll
    for (int p = 0; p < N; p++) nmax += sc[p] == smax;
    if (nmax == 1) {   // judge :308-309
#pragma unroll
        for (int p = 0; p < N; p++) out[p] = sc[p] == smax ? 1.f : -1.f;
        return true;
    }
    const int sentinel = (r.flags & SPL_F_REFCOMPAT) ? (int)(int8_t)999 : 127;   // int8(999) == -25 :313
    int mk[N], mn = 127, cnt = 0;
#pragma unroll
    for (int p = 0; p < N; p++) {
        int nc = (int)(int8_t)spl_sum5(s, L::PCARDS + p);
        mk[p] = sc[p] < smax ? sentinel : nc;
        mn = mk[p] < mn ? mk[p] : mn;
    }
#pragma unroll
    for (int p = 0; p < N; p++) cnt += mk[p] == mn;
#pragma unroll
    for (int p = 0; p < N; p++) out[p] = mk[p] == mn ? (cnt > 1 ? 0.01f : 1.f) : -1.f;
    return true;
}

// ------------------------------------------------------------------------------------------
// canonical rotation: new[i] = old[(i + shift) % size] on the four per-player blocks (:338-347)
// ------------------------------------------------------------------------------------------
// new[i] = old[(i + SHIFT) % SIZE] in place by cycle-following ("juggling": gcd(SIZE, SHIFT) cycles, each row moved
// once, one row of 7 cells held in registers). Rolled loops on purpose: the unrolled form was 300+ straight-line
// instructions that every warp fetched every ply, and these kernels are bound by instruction fetch.
template <int A, int B> struct SplGcd { static constexpr int value = SplGcd<B, A % B>::value; };
template <int A> struct SplGcd<A, 0> { static constexpr int value = A; };
template <int SIZE, int SHIFT, class S>
SPL_D void spl_roll_block(S& s, int row0) {
    constexpr int G = SplGcd<SIZE, SHIFT>::value;
#pragma unroll 1
    for (int start = 0; start < G; start++) {
        int saved[7];
#pragma unroll
        for (int c = 0; c < 7; c++) saved[c] = s.get(row0 + start, c);
        int j = start;
#pragma unroll 1
        for (;;) {
            int nxt = j + SHIFT;
            nxt = nxt >= SIZE ? nxt - SIZE : nxt;
            if (nxt == start) break;
#pragma unroll
            for (int c = 0; c < 7; c++) s.set(row0 + j, c, s.get(row0 + nxt, c));
            j = nxt;
        }
#pragma unroll
        for (int c = 0; c < 7; c++) s.set(row0 + j, c, saved[c]);
    }
}

template <int N, class S>
SPL_D void spl_rotate(S& s, int k, SplRules r) {
    typedef SplLay<N> L;
    for (int it = 0; it < k; it++) {
        spl_roll_block<N, 1>(s, L::PGEMS);
        if ((r.flags & SPL_F_REFCOMPAT) || N == 2) spl_roll_block<N * (N + 1), 3 % (N * (N + 1))>(s, L::PNOBLES);   // :345 (F7a)
        else spl_roll_block<N * (N + 1), N + 1>(s, L::PNOBLES);
        spl_roll_block<N, 1>(s, L::PCARDS);
        spl_roll_block<6 * N, 6>(s, L::PRES);
    }
}

// ------------------------------------------------------------------------------------------
// game start
// ------------------------------------------------------------------------------------------
// value of cell (row, col) in the start position before any chance event (init_game :222-233)
template <int N>
SPL_D int spl_init_cell(int row, int col) {
    typedef SplLay<N> L;
    if (row == L::BANK) return col < 5 ? L::GEMS0 : (col == SPL_GOLD ? 5 : 0);
    if (row >= L::DECK && row < L::DECK + 6 && col < 5) {
        const int t = (row - L::DECK) >> 1, k = t == 0 ? 8 : (t == 1 ? 6 : 4);
        return ((row - L::DECK) & 1) ? (int)(int8_t)(uint8_t)(0xFF00u >> k) : k;
    }
    return 0;
}
template <class S>
SPL_D void spl_write_noble(S& s, int row, int noble_id) {
    uint32_t pk = SPL_NOBLES[noble_id];
#pragma unroll
    for (int c = 0; c < 5; c++) s.set(row, c, (int)((pk >> (4 * c)) & 15u));
    s.set(row, 5, 0);
    s.set(row, 6, 3);
}

// the noble draw of init_game (:241): n+1 distinct ids of 10 by a partial Fisher-Yates over a nibble-packed permutation,
// random words w[0..4] = Philox (counter 0, stream 3) words 0..3 then (counter 1, stream 3) word 0
template <int N, class S>
SPL_D void spl_init_nobles(S& s, const uint32_t* w) {
    typedef SplLay<N> L;
    uint64_t perm = 0x9876543210ull;
#pragma unroll
    for (int i = 0; i < L::NUM_NOBLES; i++) {
        int j = i + (int)SPL_MULHI(w[i], (uint32_t)(10 - i));
        uint32_t vi = (uint32_t)(perm >> (4 * i)) & 15u, vj = (uint32_t)(perm >> (4 * j)) & 15u;
        perm &= ~((15ull << (4 * i)) | (15ull << (4 * j)));
        perm |= ((uint64_t)vj << (4 * i)) | ((uint64_t)vi << (4 * j));
        spl_write_noble(s, L::NOBLES + i, (int)vj);
    }
}

template <int N, class S>
SPL_D void spl_init_empty(S& s) {   // init_game :222-233 (no chance)
    typedef SplLay<N> L;
    for (int row = 0; row < L::ROWS; row++)
#pragma unroll
        for (int c = 0; c < 7; c++) s.set(row, c, 0);
#pragma unroll
    for (int c = 0; c < 5; c++) s.set(L::BANK, c, L::GEMS0);
    s.set(L::BANK, SPL_GOLD, 5);
#pragma unroll
    for (int t = 0; t < 3; t++) {
        const int k = t == 0 ? 8 : (t == 1 ? 6 : 4);
#pragma unroll
        for (int c = 0; c < 5; c++) {
            s.set(L::DECK + 2 * t, c, k);
            s.set(L::DECK + 2 * t + 1, c, (int)(int8_t)(uint8_t)(0xFF00u >> k));   // my_packbits(ones(k)) :233
        }
    }
}


// explicit start: deals[12] = colour*8+idx per visible slot, nobles[N+1] = noble ids (replay of a reference game)
template <int N, class S>
SPL_D void spl_init_explicit(S& s, const uint8_t* deals, const uint8_t* nobles) {
    typedef SplLay<N> L;
    spl_init_empty<N>(s);
    for (int slot = 0; slot < 12; slot++) spl_write_card(s, L::CARDS + 2 * slot, spl_deck_take<N>(s, slot >> 2, deals[slot]));
    for (int i = 0; i < L::NUM_NOBLES; i++) spl_write_noble(s, L::NOBLES + i, nobles[i]);
}

template <int N, class S>
SPL_COLD void spl_init_philox(S& s, uint64_t seed, uint32_t game, uint32_t episode) {
    typedef SplLay<N> L;
    spl_init_empty<N>(s);
    for (int slot = 0; slot < 12; slot++) {   // :237-239
        int code = spl_draw_philox<N>(s, slot >> 2, seed, game, episode, (uint32_t)slot, 2);
        spl_write_card(s, L::CARDS + 2 * slot, spl_deck_take<N>(s, slot >> 2, code));
    }
    // n+1 distinct nobles of 10 (:241)
    const SplPhilox w0 = spl_philox(seed, game, episode, 0, 3), w1 = spl_philox(seed, game, episode, 1, 3);
    const uint32_t w[5] = {w0.v[0], w0.v[1], w0.v[2], w0.v[3], w1.v[0]};
    spl_init_nobles<N>(s, w);
}

// uniform pick among the set bits of a mask (Philox stream 1): the bench's rollout policy
SPL_D int spl_pick_random(const uint32_t* m, uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply) {
    int cnt = 0;
#pragma unroll
    for (int w = 0; w < SPL_MASK_WORDS; w++) cnt += SPL_POPC(m[w]);
    if (cnt == 0) return -1;
    SplPhilox ph = spl_philox(seed, game, episode, ply, 1);
    int k = (int)SPL_MULHI(ph.v[0], (uint32_t)cnt);
    uint32_t x = 0u;   // the word that holds the k-th (0-based) set bit, found by one branch-free scan ...
    int base = 0;
    bool found = false;
#pragma unroll
    for (int w = 0; w < SPL_MASK_WORDS; w++) {
        const int pc = SPL_POPC(m[w]);
        const bool take = !found && k < pc;
        x = take ? m[w] : x;
        base = take ? 32 * w : base;
        found = found || take;
        k = found ? k : k - pc;
    }
    int bitpos = 0;    // ... then its position inside the word by a popcount binary search (no data-dependent loops)
#pragma unroll
    for (int sh = 16; sh >= 1; sh >>= 1) {
        const int c = SPL_POPC((x >> bitpos) & ((1u << sh) - 1u));
        if (k >= c) { k -= c; bitpos += sh; }
    }
    return base + bitpos;
}
