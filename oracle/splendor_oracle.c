/* TEST INFRASTRUCTURE - see splendor_oracle.h. Plain C restatement of the reference rules.
 * Every function cites the reference lines (SplendorLogicNumba.py unless noted) it follows.
 * Style on purpose mirrors the reference's array semantics (int8 wrap-around included) rather
 * than the bit-parallel form the CUDA kernels use, so the two are independent derivations. */
#include "splendor_oracle.h"
#include <string.h>
#include <stdlib.h>
#include "spo_tables.inc"

enum { C_GOLD = 5, C_PTS = 6 };

/* row offsets of the eight views bound by copy_state :291-303 */
typedef struct { int bank, cards, deck, nobles, pgems, pnobles, pcards, pres, rows, num_nobles; } views;
static views vw(int n) {
    views v;
    v.bank = 0; v.cards = 1; v.deck = 25; v.nobles = 31;
    v.pgems = 32 + n; v.pnobles = 32 + 2 * n; v.pcards = 32 + 3 * n + n * n; v.pres = 32 + 4 * n + n * n;
    v.rows = 32 + 10 * n + n * n; v.num_nobles = n + 1;
    return v;
}
#define ROW(st, r) ((st) + 7 * (r))
#define CROW(st, r) ((const int8_t*)((st) + 7 * (r)))

int spo_rows(int n) { return 32 + 10 * n + n * n; }

void spo_default_rules(spo_rules* r, int n) {
    r->n_players = n; r->token_limit = 10; r->enable_reserve = 1; r->enable_giveback = 1; r->ref_compat = 1;
}

static int sum5(const int8_t* row) { return row[0] + row[1] + row[2] + row[3] + row[4]; }
static int sum7(const int8_t* row) { return sum5(row) + row[5] + row[6]; }

/* ------------------------------------------------------------------ init :222-246 */
void spo_init_empty(int8_t* st, const spo_rules* r) {
    int n = r->n_players; views v = vw(n);
    memset(st, 0, (size_t)7 * v.rows);
    int gems = (n == 2) ? 4 : (n == 3) ? 5 : 7;          /* num_gems_in_play :90 */
    int8_t* bank = ROW(st, v.bank);
    for (int c = 0; c < 5; c++) bank[c] = (int8_t)gems;
    bank[C_GOLD] = 5; bank[C_PTS] = 0;
    for (int t = 0; t < 3; t++) {
        int k = SPO_NCARDS[t];
        for (int c = 0; c < 5; c++) {
            ROW(st, v.deck + 2 * t)[c] = (int8_t)k;                       /* :231 */
            /* my_packbits(ones(k)) with mask [128..1] :23,:43-46 -> top k bits, stored as int8 */
            ROW(st, v.deck + 2 * t + 1)[c] = (int8_t)(uint8_t)(0xFF00 >> k);
        }
    }
}

static int deck_take(int8_t* st, int tier, int color, int idx) {  /* _get_select_card :423-443 */
    views v = vw(2);
    uint8_t bits = (uint8_t)ROW(st, v.deck + 2 * tier + 1)[color];
    uint8_t m = (uint8_t)(128u >> idx);
    if (!(bits & m)) return -1;
    ROW(st, v.deck + 2 * tier + 1)[color] = (int8_t)(uint8_t)(bits & ~m);
    ROW(st, v.deck + 2 * tier)[color] -= 1;
    return 0;
}

int spo_deal_to_slot(int8_t* st, const spo_rules* r, int slot, int color, int idx) {
    (void)r;
    int tier = slot / 4, index = slot % 4;
    if (deck_take(st, tier, color, idx)) return -1;
    memcpy(ROW(st, 1 + 8 * tier + 2 * index), SPO_CARDS[tier][color][idx][0], 7);   /* :450 */
    memcpy(ROW(st, 1 + 8 * tier + 2 * index + 1), SPO_CARDS[tier][color][idx][1], 7);
    return 0;
}

void spo_set_noble(int8_t* st, const spo_rules* r, int slot, int noble_id) {
    views v = vw(r->n_players);
    memcpy(ROW(st, v.nobles + slot), SPO_NOBLES[noble_id], 7);                      /* :243 */
}

/* ------------------------------------------------------------------ legality :251-265, :476-680 */
static void valid_buy_rows(const int8_t* st, const views* v, int player, const int* cost_rows, int k, uint8_t* out) {
    /* _valid_buy :476-501 and _valid_buy_reserve :538-552 share this body */
    const int8_t* g = CROW(st, v->pgems + player);
    const int8_t* pc = CROW(st, v->pcards + player);
    for (int i = 0; i < k; i++) {
        const int8_t* cost = CROW(st, cost_rows[i]);
        int missing = 0, s = 0;
        for (int c = 0; c < 5; c++) {
            int8_t d = (int8_t)(cost[c] - g[c] - pc[c]);      /* int8 arithmetic as numpy does */
            if (d > 0) missing += d;
            s += cost[c];
        }
        out[i] = (uint8_t)((missing <= g[C_GOLD]) && (s != 0));
    }
}

static void valid_reserve(const int8_t* st, const spo_rules* r, const views* v, int player, int is_limit, uint8_t* out15) {
    /* _valid_reserve :508-515 */
    const int8_t* bank = CROW(st, v->bank);
    if (((!r->enable_reserve) || (sum7(CROW(st, v->pgems + player)) == r->token_limit && bank[C_GOLD] > 0)) && is_limit) {
        memset(out15, 0, 15);
        return;
    }
    int empty_slot = (sum5(CROW(st, v->pres + 6 * player + 5)) == 0);   /* gain row of the 3rd slot :514 */
    for (int i = 0; i < 12; i++) out15[i] = (uint8_t)((sum5(CROW(st, v->cards + 2 * i)) != 0) && empty_slot);
    for (int t = 0; t < 3; t++) out15[12 + t] = (uint8_t)((sum5(CROW(st, v->deck + 2 * t)) != 0) && empty_slot);
}

static void valid_get_gems(const int8_t* st, const spo_rules* r, const views* v, int player, int is_limit, uint8_t* out25) {
    /* _valid_get_gems :562-576 */
    const int8_t* bank = CROW(st, v->bank);
    int npg = sum7(CROW(st, v->pgems + player));
    int nspec = 0;
    for (int c = 0; c < 5; c++) nspec += (bank[c] != 0);
    for (int i = 0; i < 25; i++) {
        int enough = 1, k = 0;
        for (int c = 0; c < 5; c++) { if ((int8_t)(bank[c] - SPO_COMB3[i][c]) < 0) enough = 0; k += SPO_COMB3[i][c]; }
        int ok = is_limit ? (enough && (npg + k <= r->token_limit)) : enough;
        out25[i] = (uint8_t)ok;
    }
    if (is_limit) {
        if (npg != 9 && nspec != 1) memset(out25, 0, 5);            /* :571-572 */
        if (npg != 8 && nspec != 2) memset(out25 + 5, 0, 10);       /* :573-574 */
    }
}

static void valid_get_identical(const int8_t* st, const spo_rules* r, const views* v, int player, int is_limit, uint8_t* out5) {
    /* _valid_get_gems_identical :578-583 */
    const int8_t* bank = CROW(st, v->bank);
    int room = (sum7(CROW(st, v->pgems + player)) + 2 <= r->token_limit);
    for (int c = 0; c < 5; c++) out5[c] = (uint8_t)((bank[c] >= 4) && (is_limit ? room : 1));
}

void spo_valid_moves(const int8_t* st, const spo_rules* r, int player, uint8_t* out) {
    int n = r->n_players; views v = vw(n);
    memset(out, 0, SPO_ACTIONS);
    int rows12[12], rows3[3];
    for (int i = 0; i < 12; i++) rows12[i] = v.cards + 2 * i;
    for (int i = 0; i < 3; i++) rows3[i] = v.pres + 6 * player + 2 * i;
    valid_buy_rows(st, &v, player, rows12, 12, out + 0);                 /* :253 */
    valid_reserve(st, r, &v, player, 1, out + 12);                       /* :254 */
    valid_buy_rows(st, &v, player, rows3, 3, out + 27);                  /* :255 */
    valid_get_gems(st, r, &v, player, 1, out + 30);                      /* :256 */
    valid_get_identical(st, r, &v, player, 1, out + 55);

    uint8_t get_flgs[30], giv[20], giv3[40], rsv[15];
    valid_get_gems(st, r, &v, player, 0, get_flgs);                      /* :257 */
    valid_get_identical(st, r, &v, player, 0, get_flgs + 25);
    const int8_t* g = CROW(st, v.pgems + player);
    for (int i = 0; i < 15; i++) {                                       /* _valid_give_gems :595-600 */
        int ok = 1;
        for (int c = 0; c < 5; c++) if ((int8_t)(g[c] - SPO_COMB2[i][c]) < 0) ok = 0;
        giv[i] = (uint8_t)(ok && r->enable_giveback);
    }
    for (int c = 0; c < 5; c++) giv[15 + c] = (uint8_t)((g[c] >= 2) && r->enable_giveback);   /* :609-613 */
    for (int i = 0; i < 40; i++) {                                       /* _valid_give_gems3 :602-607 */
        int ok = 1;
        for (int c = 0; c < 5; c++) if ((int8_t)(g[c] - SPO_GIVE3[i][c]) < 0) ok = 0;
        giv3[i] = (uint8_t)(ok && r->enable_giveback);
    }
    valid_reserve(st, r, &v, player, 0, rsv);                            /* :260 */

    /* _valid_exchange :615-680 ; output block starts at 60 */
    uint8_t* ex = out + 60;
    int tokens = sum7(g);
    if (tokens > 7) {
        const uint8_t* same2 = get_flgs + 25; const uint8_t* dif2 = get_flgs + 5; const uint8_t* dif3 = get_flgs + 15;
        uint8_t* t3g1 = ex;        uint8_t* t3g2 = ex + 20;   uint8_t* t2dg2 = ex + 50;  uint8_t* t2sg2 = ex + 110;
        uint8_t* t2dg1 = ex + 160; uint8_t* t2sg1 = ex + 190; uint8_t* t1g1 = ex + 210;  uint8_t* t1gg1 = ex + 230;
        uint8_t* t3g3 = ex + 305;
        if (tokens == r->token_limit - 2) {                              /* :632-636 */
            for (int i = 0; i < 10; i++) for (int j = 0; j < 2; j++) t3g1[2 * i + j] = dif3[i] && giv[SPO_GIVE_IDS[0][i][j]];
        } else if (tokens == r->token_limit - 1) {                       /* :638-652 */
            for (int i = 0; i < 10; i++) for (int j = 0; j < 3; j++) t3g2[3 * i + j] = dif3[i] && giv[SPO_GIVE_IDS[1][i][j]];
            for (int i = 0; i < 10; i++) for (int j = 0; j < 3; j++) t2dg1[3 * i + j] = dif2[i] && giv[SPO_GIVE_IDS[4][i][j]];
            for (int i = 0; i < 5; i++) for (int j = 0; j < 4; j++) t2sg1[4 * i + j] = same2[i] && giv[SPO_GIVE_IDS[5][i][j]];
        } else {                                                          /* :654-678 */
            for (int i = 0; i < 10; i++) for (int j = 0; j < 6; j++) t2dg2[6 * i + j] = dif2[i] && giv[SPO_GIVE_IDS[2][i][j]];
            for (int i = 0; i < 5; i++) for (int j = 0; j < 10; j++) t2sg2[10 * i + j] = same2[i] && giv[SPO_GIVE_IDS[3][i][j]];
            for (int k = 0; k < 20; k++) t1g1[k] = get_flgs[k / 4] && giv[SPO_GIVE1_TAKE1[k]];
            for (int k = 0; k < 40; k++) t3g3[k] = dif3[k / 4] && giv3[k];
            if (CROW(st, v.bank)[C_GOLD] > 0)
                for (int k = 0; k < 75; k++) t1gg1[k] = rsv[k / 5] && giv[k % 5];
        }
    }
    int any = 0;
    for (int i = 0; i < SPO_ACTIONS - 1; i++) any |= out[i];
    out[SPO_ACTIONS - 1] = (uint8_t)!any;                                /* :263 */
}

/* ------------------------------------------------------------------ Philox4x32-10 (Salmon et al. 2011) */
void spo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; i++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

static void philox(uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply, uint32_t stream, uint32_t out[4]) {
    uint32_t ctr[4] = { game, episode, ply, stream };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    spo_philox4x32_10(ctr, key, out);
}

/* Same two-stage draw as _get_deck_card :400-412 (colour in proportion to the remaining count,
 * then uniform among that colour's remaining cards), with the two float64 uniforms replaced by
 * 32-bit Philox words: searchsorted(cumsum(p), u, 'right') == first c with cumsum_c > floor(u*total). */
int spo_philox_draw(const int8_t* st, int tier, uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply, uint32_t stream) {
    const int8_t* cnt = CROW(st, 25 + 2 * tier);
    const int8_t* bitsrow = CROW(st, 25 + 2 * tier + 1);
    int total = sum5(cnt);
    if (total == 0) return -1;
    uint32_t w[4];
    philox(seed, game, episode, ply, stream, w);
    int k = (int)mulhi32(w[0], (uint32_t)total), color = 0, acc = 0;
    for (color = 0; color < 5; color++) { acc += cnt[color]; if (acc > k) break; }
    uint8_t bits = (uint8_t)bitsrow[color];
    int j = (int)mulhi32(w[1], (uint32_t)cnt[color]);
    for (int idx = 0; idx < 8; idx++)
        if (bits & (128u >> idx)) { if (j == 0) return color * 8 + idx; j--; }
    return -1;
}

void spo_init_philox(int8_t* st, const spo_rules* r, uint64_t seed, uint32_t game, uint32_t episode) {
    spo_init_empty(st, r);
    for (int slot = 0; slot < 12; slot++) {                                    /* :237-239 */
        int d = spo_philox_draw(st, slot / 4, seed, game, episode, (uint32_t)slot, 2);
        spo_deal_to_slot(st, r, slot, d / 8, d % 8);
    }
    /* nobles: n+1 distinct of 10 (:241), partial Fisher-Yates on Philox stream 3 */
    uint32_t w[8];
    philox(seed, game, episode, 0, 3, w);
    philox(seed, game, episode, 1, 3, w + 4);
    int perm[10];
    for (int i = 0; i < 10; i++) perm[i] = i;
    for (int i = 0; i <= r->n_players; i++) {
        int j = i + (int)mulhi32(w[i], (uint32_t)(10 - i));
        int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
        spo_set_noble(st, r, i, perm[i]);
    }
}

int spo_philox_pick(const uint8_t* valids, uint64_t seed, uint32_t game, uint32_t episode, uint32_t ply) {
    int cnt = 0;
    for (int i = 0; i < SPO_ACTIONS; i++) cnt += valids[i] != 0;
    if (cnt == 0) return -1;
    uint32_t w[4];
    philox(seed, game, episode, ply, 1, w);
    int k = (int)mulhi32(w[0], (uint32_t)cnt);
    for (int i = 0; i < SPO_ACTIONS; i++) if (valids[i]) { if (k == 0) return i; k--; }
    return -1;
}

/* ------------------------------------------------------------------ moves :267-289, :445-474, :503-560, :585-593, :685-768 */
typedef struct { int mode; int color, idx; uint64_t seed; uint32_t game, episode, ply; } chance;  /* mode: 0 det, 1 replay, 2 philox */

static int draw_card(int8_t* st, int tier, const chance* ch, int* color, int* idx) {
    /* _get_deck_card :400-420 with the outcome supplied (replay) or Philox-drawn */
    if (sum5(CROW(st, 25 + 2 * tier)) == 0) return 0;        /* no more cards :402 */
    if (ch->mode == 1) { *color = ch->color; *idx = ch->idx; }
    else {
        int d = spo_philox_draw(st, tier, ch->seed, ch->game, ch->episode, ch->ply, 0);
        *color = d / 8; *idx = d % 8;
    }
    if (deck_take(st, tier, *color, *idx)) return -1;
    return 1;
}

static int fill_new_card(int8_t* st, int tier, int index, const chance* ch) {   /* _fill_new_card :445-450 */
    int8_t* slot = ROW(st, 1 + 8 * tier + 2 * index);
    memset(slot, 0, 14);
    if (ch->mode != 0) {
        int color, idx, k = draw_card(st, tier, ch, &color, &idx);
        if (k < 0) return -1;
        if (k == 1) { memcpy(slot, SPO_CARDS[tier][color][idx][0], 7); memcpy(slot + 7, SPO_CARDS[tier][color][idx][1], 7); }
    }
    return 0;
}

static void give_nobles_if_earned(int8_t* st, const views* v, int player) {      /* :763-768 */
    const int8_t* pc = CROW(st, v->pcards + player);
    for (int i = 0; i < v->num_nobles; i++) {
        int8_t* nb = ROW(st, v->nobles + i);
        int ok = sum5(nb) > 0;
        for (int c = 0; c < 5; c++) if (pc[c] < nb[c]) ok = 0;
        if (ok) {
            memcpy(ROW(st, v->pnobles + v->num_nobles * player + i), nb, 7);      /* writer stride = n+1 :767 */
            memset(nb, 0, 7);
        }
    }
}

static void buy_card(int8_t* st, const views* v, int cost_row, int player) {     /* _buy_card :458-474 (P4 applied) */
    int8_t* cost = ROW(st, cost_row); int8_t* gain = cost + 7;
    int8_t* g = ROW(st, v->pgems + player); int8_t* pc = ROW(st, v->pcards + player); int8_t* bank = ROW(st, v->bank);
    int missing = 0; int8_t paid[5];
    for (int c = 0; c < 5; c++) {
        int8_t d = (int8_t)(cost[c] - g[c] - pc[c]);
        if (d > 0) missing += d;
        int8_t need = (int8_t)(cost[c] - pc[c]); if (need < 0) need = 0;
        paid[c] = need < g[c] ? need : g[c];
    }
    for (int c = 0; c < 5; c++) { g[c] = (int8_t)(g[c] - paid[c]); bank[c] = (int8_t)(bank[c] + paid[c]); }
    g[C_GOLD] = (int8_t)(g[C_GOLD] - missing);
    bank[C_GOLD] = (int8_t)(bank[C_GOLD] + missing);
    for (int c = 0; c < 7; c++) pc[c] = (int8_t)(pc[c] + gain[c]);               /* :472 */
    give_nobles_if_earned(st, v, player);
}

static void get_gems(int8_t* st, const views* v, int i, int player) {            /* _get_gems :585-593 */
    int8_t gems[5] = {0, 0, 0, 0, 0};
    if (i < 25) memcpy(gems, SPO_COMB3[i], 5); else gems[i - 25] = 2;
    for (int c = 0; c < 5; c++) { ROW(st, v->bank)[c] -= gems[c]; ROW(st, v->pgems + player)[c] += gems[c]; }
}

static void give_gems(int8_t* st, const views* v, int i, int player) {           /* _give_gems :685-694 */
    int8_t gems[5] = {0, 0, 0, 0, 0};
    if (i < 15) memcpy(gems, SPO_COMB2[i], 5); else gems[i - 15] = 2;
    for (int c = 0; c < 5; c++) { ROW(st, v->bank)[c] += gems[c]; ROW(st, v->pgems + player)[c] -= gems[c]; }
}

static int reserve(int8_t* st, const views* v, int i, int player, const chance* ch) {   /* _reserve :517-536 */
    int empty_slot = -1;
    for (int s = 0; s < 3; s++) {
        int row = v->pres + 6 * player + 2 * s;
        if (sum5(CROW(st, row)) == 0) { empty_slot = row; break; }               /* first free COST row :521 */
    }
    if (empty_slot < 0) return -2;                                               /* reference: unbound local */
    if (i < 12) {
        int tier = i / 4, index = i % 4;
        memcpy(ROW(st, empty_slot), ROW(st, 1 + 8 * tier + 2 * index), 14);
        if (fill_new_card(st, tier, index, ch)) return -1;
    } else if (ch->mode != 0) {                                                   /* deck reserve; deterministic: nothing stored :529-532 */
        int tier = i - 12, color, idx, k = draw_card(st, tier, ch, &color, &idx);
        if (k <= 0) return k < 0 ? -1 : -3;                                       /* reference would assign None */
        memcpy(ROW(st, empty_slot), SPO_CARDS[tier][color][idx][0], 7);
        memcpy(ROW(st, empty_slot + 1), SPO_CARDS[tier][color][idx][1], 7);
    }
    if (ROW(st, v->bank)[C_GOLD] > 0) {                                           /* :534-536 */
        ROW(st, v->pgems + player)[C_GOLD] += 1;
        ROW(st, v->bank)[C_GOLD] -= 1;
    }
    return 0;
}

static void give_and_get(int8_t* st, const views* v, int i, int player) {        /* _give_and_get_gems :697-756 */
    if (i < 20)       { int t = i / 2;  get_gems(st, v, t + 15, player); give_gems(st, v, SPO_GIVE_IDS[0][t][i % 2], player); }
    else if (i < 50)  { i -= 20;  int t = i / 3;  get_gems(st, v, t + 15, player); give_gems(st, v, SPO_GIVE_IDS[1][t][i % 3], player); }
    else if (i < 110) { i -= 50;  int t = i / 6;  get_gems(st, v, t + 5, player);  give_gems(st, v, SPO_GIVE_IDS[2][t][i % 6], player); }
    else if (i < 160) { i -= 110; int t = i / 10; get_gems(st, v, t + 25, player); give_gems(st, v, SPO_GIVE_IDS[3][t][i % 10], player); }
    else if (i < 190) { i -= 160; int t = i / 3;  get_gems(st, v, t + 5, player);  give_gems(st, v, SPO_GIVE_IDS[4][t][i % 3], player); }
    else if (i < 210) { i -= 190; int t = i / 4;  get_gems(st, v, t + 25, player); give_gems(st, v, SPO_GIVE_IDS[5][t][i % 4], player); }
    else if (i < 230) { i -= 210; get_gems(st, v, i / 4, player); give_gems(st, v, SPO_GIVE1_TAKE1[i], player); }
    else { i -= 305; get_gems(st, v, SPO_GIVE_IDS3[i][0] + 15, player);
           give_gems(st, v, SPO_GIVE_IDS3[i][1], player); give_gems(st, v, SPO_GIVE_IDS3[i][2], player); }
}

int spo_make_move(int8_t* st, const spo_rules* r, int move, int player, int reveal, uint64_t seed, uint32_t game, uint32_t episode) {
    int n = r->n_players; views v = vw(n);
    chance ch; memset(&ch, 0, sizeof ch);
    if (reveal == -1) ch.mode = 0;
    else if (reveal == -2) { ch.mode = 2; ch.seed = seed; ch.game = game; ch.episode = episode; ch.ply = (uint32_t)spo_get_round(st); }
    else { ch.mode = 1; ch.color = reveal / 8; ch.idx = reveal % 8; }
    int rc = 0;
    if (move < 12) {                                                              /* _buy :503-506 */
        buy_card(st, &v, v.cards + 2 * move, player);
        rc = fill_new_card(st, move / 4, move % 4, &ch);
    } else if (move < 27) {
        rc = reserve(st, &v, move - 12, player, &ch);
    } else if (move < 30) {                                                       /* _buy_reserve :554-560 */
        int i = move - 27, start = v.pres + 6 * player + 2 * i;
        buy_card(st, &v, start, player);
        if (i < 2) memmove(ROW(st, start), ROW(st, start + 2), (size_t)7 * (v.pres + 6 * player + 4 - start));
        memset(ROW(st, v.pres + 6 * player + 4), 0, 14);
    } else if (move < 60) {
        get_gems(st, &v, move - 30, player);
    } else if (move < 290) {
        give_and_get(st, &v, move - 60, player);
    } else if (move < 365) {                                                      /* _reserve_and_give :759-761 */
        rc = reserve(st, &v, (move - 290) / 5, player, &ch);
        if (rc == 0) give_gems(st, &v, (move - 290) % 5, player);
    } else if (move < 405) {
        give_and_get(st, &v, move - 60, player);
    } /* 405: pass = no-op (patch P6) */
    if (rc < 0) return rc;
    ROW(st, v.bank)[C_PTS] += 1;                                                  /* ply counter :287 */
    return (player + 1) % n;
}

/* ------------------------------------------------------------------ end of game :217-220, :306-334, :397 */
int spo_get_round(const int8_t* st) { return (uint8_t)st[C_PTS]; }

int spo_get_score(const int8_t* st, const spo_rules* r, int player) {
    views v = vw(r->n_players);
    int stride = r->ref_compat ? 3 : v.num_nobles;                                /* reader stride :219 (F7a) */
    int pts = CROW(st, v.pcards + player)[C_PTS];
    for (int i = 0; i < stride; i++) pts += CROW(st, v.pnobles + stride * player + i)[C_PTS];
    return pts;
}

void spo_check_end_game(const int8_t* st, const spo_rules* r, float* out) {
    int n = r->n_players; views v = vw(n);
    for (int p = 0; p < n; p++) out[p] = 0.f;
    int round = spo_get_round(st);
    if (round % n != 0) return;                                                   /* :322 */
    int8_t scores[4], ncards[4];
    int smax = -128;
    for (int p = 0; p < n; p++) { scores[p] = (int8_t)spo_get_score(st, r, p); if (scores[p] > smax) smax = scores[p]; }
    int max_moves = (uint8_t)(62 * n);                                            /* uint8 field :66,:92 */
    if (!(smax >= 15 || round >= max_moves)) return;
    int nmax = 0;
    for (int p = 0; p < n; p++) nmax += (scores[p] == smax);
    for (int p = 0; p < n; p++) ncards[p] = (int8_t)sum5(CROW(st, v.pcards + p));
    if (nmax == 1) {                                                              /* judge :308-309 */
        for (int p = 0; p < n; p++) out[p] = (scores[p] == smax) ? 1.f : -1.f;
        return;
    }
    int8_t sentinel = r->ref_compat ? (int8_t)999 : (int8_t)127;                  /* int8(999) == -25 :313 (F7b) */
    int8_t masked[4]; int mn = 127;
    for (int p = 0; p < n; p++) { masked[p] = (scores[p] < smax) ? sentinel : ncards[p]; if (masked[p] < mn) mn = masked[p]; }
    int cnt = 0;
    for (int p = 0; p < n; p++) cnt += (masked[p] == mn);
    for (int p = 0; p < n; p++) out[p] = (masked[p] == mn) ? (cnt > 1 ? 0.01f : 1.f) : -1.f;
}

/* ------------------------------------------------------------------ canonical rotation :338-347 */
static void roll_rows(int8_t* base, int size, int shift) {
    int8_t tmp[7 * 64];
    memcpy(tmp, base, (size_t)7 * size);
    for (int i = 0; i < size; i++) memcpy(base + 7 * i, tmp + 7 * ((i + shift) % size), 7);
}

void spo_swap_players(int8_t* st, const spo_rules* r, int k) {
    int n = r->n_players; views v = vw(n);
    int nstride = r->ref_compat ? 3 : v.num_nobles;                               /* :345 (F7a) */
    roll_rows(ROW(st, v.pgems), n, 1 * k);
    roll_rows(ROW(st, v.pnobles), n * v.num_nobles, nstride * k);
    roll_rows(ROW(st, v.pcards), n, 1 * k);
    roll_rows(ROW(st, v.pres), 6 * n, 6 * k);
}

/* ------------------------------------------------------------------ symmetries :349-395 */
static int nb_reserved(const int8_t* st, const views* v, int player) {           /* :770-774 */
    for (int c = 0; c < 3; c++) if (sum5(CROW(st, v->pres + 6 * player + 2 * c)) == 0) return c;
    return 3;
}

int spo_symmetries(const int8_t* st, const spo_rules* r, const float* pi, const uint8_t* valids,
                   int8_t* out_states, float* out_pi, uint8_t* out_valids) {
    int n = r->n_players; views v = vw(n); size_t S = (size_t)7 * v.rows; int k = 0;
#define EMIT_BEGIN() int8_t* os = out_states + S * k; float* op = out_pi + (size_t)SPO_ACTIONS * k; uint8_t* ov = out_valids + (size_t)SPO_ACTIONS * k; \
                     memcpy(os, st, S); memcpy(op, pi, sizeof(float) * SPO_ACTIONS); memcpy(ov, valids, SPO_ACTIONS)
    { EMIT_BEGIN(); (void)os; (void)op; (void)ov; k++; }                          /* identity :367 */
    for (int tier = 0; tier < 3; tier++)
        for (int s = 0; s < 3; s++) {                                             /* :369-376 */
            const int8_t* perm = SPO_CARD_SYM[s];
            EMIT_BEGIN();
            for (int j = 0; j < 4; j++) {
                memcpy(ROW(os, 1 + 8 * tier + 2 * j), CROW(st, 1 + 8 * tier + 2 * perm[j]), 14);
                op[4 * tier + j] = pi[4 * tier + perm[j]];           op[12 + 4 * tier + j] = pi[12 + 4 * tier + perm[j]];
                ov[4 * tier + j] = valids[4 * tier + perm[j]];       ov[12 + 4 * tier + j] = valids[12 + 4 * tier + perm[j]];
            }
            k++;
        }
    for (int player = 0; player < n; player++) {                                  /* :379-393 */
        int nb = nb_reserved(st, &v, player);
        for (int s = 0; s < 2; s++) {
            const int8_t* perm = SPO_RES_SYM[nb][s];
            if (perm[0] < 0) continue;
            EMIT_BEGIN();
            for (int j = 0; j < 3; j++) {
                memcpy(ROW(os, v.pres + 6 * player + 2 * j), CROW(st, v.pres + 6 * player + 2 * perm[j]), 14);
                if (player == 0) { op[27 + j] = pi[27 + perm[j]]; ov[27 + j] = valids[27 + perm[j]]; }
            }
            k++;
        }
    }
#undef EMIT_BEGIN
    return k;
}

/* ------------------------------------------------------------------ CPU baseline driver */
long spo_rollout(const spo_rules* r, uint64_t seed, uint32_t game0, int games, int32_t* plies_out, float* result_out) {
    int n = r->n_players; long total = 0;
    int8_t st[7 * 96]; uint8_t valids[SPO_ACTIONS]; float res[4];
    for (int g = 0; g < games; g++) {
        uint32_t game = game0 + (uint32_t)g;
        spo_init_philox(st, r, seed, game, 0);
        int player = 0, plies = 0;
        for (;;) {
            spo_valid_moves(st, r, player, valids);
            int a = spo_philox_pick(valids, seed, game, 0, (uint32_t)spo_get_round(st));
            player = spo_make_move(st, r, a, player, -2, seed, game, 0);
            plies++;
            spo_check_end_game(st, r, res);
            int ended = 0;
            for (int p = 0; p < n; p++) ended |= (res[p] != 0.f);
            if (ended || player < 0) break;
        }
        total += plies;
        if (plies_out) plies_out[g] = plies;
        if (result_out) memcpy(result_out + (size_t)n * g, res, sizeof(float) * n);
    }
    return total;
}
