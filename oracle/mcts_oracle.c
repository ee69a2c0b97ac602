/* TEST INFRASTRUCTURE - see mcts_oracle.h. Mirrors MCTS.py statement by statement: same dictionary-of-nodes
 * semantics (exact state bytes as key), same numeric types (Ps float32, Qsa float64, Nsa int64, Qs float32
 * under NumPy >= 2 scalar rules), same evaluation order. Compiled with -ffp-contract=off. */
#include "mcts_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define A SPO_ACTIONS
#define NAN_SENTINEL (-42.0)   /* MCTS.py:9 */
#define EPS 1e-8               /* :8 */
#define KFORCED 0.5            /* :10 */

typedef struct {
    int8_t*  state;      /* key */
    int      has_es;     /* Es is not None */
    int      terminal;
    float    es[4];
    int      has_ps;     /* Ps is not None (expanded) */
    uint8_t* vs;         /* [406] */
    float*   ps;         /* [406] */
    long     ns;
    double*  qsa;        /* [406] */
    int64_t* nsa;        /* [406] */
    int      r;
    float    qs;
} node;

struct mo_tree {
    spo_rules rules; mo_args args; mo_predict_fn fn; void* user;
    int n, sbytes;
    node* nodes; long n_nodes, cap_nodes;
    long* table; long tsize;     /* open addressing: index+1 */
    long nn_calls;
    int step;
};

static uint64_t fnv(const int8_t* s, int len) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (int i = 0; i < len; i++) { h ^= (uint8_t)s[i]; h *= 0x100000001B3ull; }
    return h;
}
static uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    return x;
}

void mo_fake_predict(const int8_t* state, int sbytes, const uint8_t* valids, int n, float* ps, float* v) {
    uint64_t h = fnv(state, sbytes);
    int first = -1; long sum = 0;
    for (int a = 0; a < A; a++) {
        ps[a] = 0.f;
        if (valids[a]) {
            long w = 1 + (long)(mix64(h + (uint64_t)a * 0x9E3779B97F4A7C15ull) >> 58);
            ps[a] = (float)w; sum += w;
            if (first < 0) first = a;
        }
    }
    if (first >= 0) ps[first] += (float)(8192 - sum);
    for (int a = 0; a < A; a++) ps[a] = ps[a] / 8192.f;
    for (int p = 0; p < n; p++) v[p] = (float)((double)((long)((h >> (8 * p + 3)) & 0x7F) - 64) / 64.0);
}

mo_tree* mo_create(const spo_rules* rules, const mo_args* args, mo_predict_fn fn, void* user) {
    mo_tree* t = (mo_tree*)calloc(1, sizeof *t);
    t->rules = *rules; t->args = *args; t->fn = fn; t->user = user;
    t->n = rules->n_players; t->sbytes = 7 * spo_rows(t->n);
    t->cap_nodes = 1024; t->nodes = (node*)calloc((size_t)t->cap_nodes, sizeof(node));
    t->tsize = 4096; t->table = (long*)calloc((size_t)t->tsize, sizeof(long));
    return t;
}
static void free_nodes(mo_tree* t) {
    for (long i = 0; i < t->n_nodes; i++) { node* nd = &t->nodes[i]; free(nd->state); free(nd->vs); free(nd->ps); free(nd->qsa); free(nd->nsa); }
    t->n_nodes = 0;
    memset(t->table, 0, sizeof(long) * (size_t)t->tsize);
}
void mo_reset(mo_tree* t) { free_nodes(t); }
void mo_destroy(mo_tree* t) { free_nodes(t); free(t->nodes); free(t->table); free(t); }
long mo_num_nodes(const mo_tree* t) { return t->n_nodes; }
long mo_nn_calls(const mo_tree* t) { return t->nn_calls; }

static void table_insert(mo_tree* t, long idx) {
    uint64_t h = fnv(t->nodes[idx].state, t->sbytes);
    long i = (long)(h % (uint64_t)t->tsize);
    while (t->table[i]) i = (i + 1) % t->tsize;
    t->table[i] = idx + 1;
}
static node* lookup(mo_tree* t, const int8_t* s) {           /* nodes_data.get(s) :120 */
    uint64_t h = fnv(s, t->sbytes);
    long i = (long)(h % (uint64_t)t->tsize);
    while (t->table[i]) {
        node* nd = &t->nodes[t->table[i] - 1];
        if (memcmp(nd->state, s, (size_t)t->sbytes) == 0) return nd;
        i = (i + 1) % t->tsize;
    }
    return NULL;
}
static node* insert(mo_tree* t, const int8_t* s) {
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes *= 2; t->nodes = (node*)realloc(t->nodes, sizeof(node) * (size_t)t->cap_nodes);
        memset(t->nodes + t->n_nodes, 0, sizeof(node) * (size_t)(t->cap_nodes - t->n_nodes));
    }
    if (2 * (t->n_nodes + 1) > t->tsize) {
        t->tsize *= 4; free(t->table); t->table = (long*)calloc((size_t)t->tsize, sizeof(long));
        for (long i = 0; i < t->n_nodes; i++) table_insert(t, i);
    }
    node* nd = &t->nodes[t->n_nodes];
    memset(nd, 0, sizeof *nd);
    nd->state = (int8_t*)malloc((size_t)t->sbytes); memcpy(nd->state, s, (size_t)t->sbytes);
    table_insert(t, t->n_nodes);
    t->n_nodes++;
    return nd;
}

static void normalise(float* ps) {             /* :239-242 float32 sum then divide */
    float s = 0.f;
    for (int a = 0; a < A; a++) s += ps[a];
    for (int a = 0; a < A; a++) ps[a] = ps[a] / s;
}
static void root_noise(mo_tree* t, float* ps, const uint8_t* vs, const double* dir) {
    if (t->args.temperature0 != 1.0) {          /* softmax :244-250 */
        for (int a = 0; a < A; a++) ps[a] = powf(ps[a], (float)(1.0 / t->args.temperature0));
        normalise(ps);
    }
    int k = 0;                                  /* applyDirNoise :180-186 */
    for (int a = 0; a < A; a++)
        if (vs[a]) { ps[a] = (float)((double)(0.75f * ps[a]) + 0.25 * dir[k]); k++; }
    normalise(ps);
}

static int pick_highest_ucb(const node* nd, double cpuct, int forced, int n_iter, double fpu) {   /* :199-219 */
    double cur_best = -INFINITY; int best = -1;
    double fpu_init = fpu > 0 ? (double)nd->qs - fpu : fpu;
    for (int a = 0; a < A; a++) {
        if (!nd->vs[a]) continue;
        if (forced) {
            if (nd->nsa[a] < (int64_t)sqrt(KFORCED * (double)nd->ps[a] * (double)n_iter)) return a;
        }
        double u;
        if (nd->qsa[a] != NAN_SENTINEL) u = nd->qsa[a] + cpuct * (double)nd->ps[a] * sqrt((double)nd->ns) / (double)(1 + nd->nsa[a]);
        else u = fpu_init + cpuct * (double)nd->ps[a] * sqrt((double)nd->ns + EPS);
        if (u > cur_best) { cur_best = u; best = a; }
    }
    return best;
}

static void search(mo_tree* t, const int8_t* s, int noise, int forced, const double* dir, float* v_out) {
    const int n = t->n;
    node* nd = lookup(t, s);
    if (!nd) { nd = insert(t, s); nd->r = spo_get_round(s); }
    if (!nd->has_es) {                                   /* :123-129 */
        spo_check_end_game(s, &t->rules, nd->es);
        nd->has_es = 1;
        nd->terminal = 0;
        for (int p = 0; p < n; p++) nd->terminal |= (nd->es[p] != 0.f);
        if (nd->terminal) { memcpy(v_out, nd->es, sizeof(float) * (size_t)n); return; }
    } else if (nd->terminal) { memcpy(v_out, nd->es, sizeof(float) * (size_t)n); return; }   /* :130-132 */

    if (!nd->has_ps) {                                   /* first visit :134-148 */
        nd->vs = (uint8_t*)malloc(A); nd->ps = (float*)malloc(sizeof(float) * A);
        spo_valid_moves(s, &t->rules, 0, nd->vs);
        float v[4];
        if (t->fn) t->fn(s, nd->vs, nd->ps, v, t->user); else mo_fake_predict(s, t->sbytes, nd->vs, n, nd->ps, v);
        t->nn_calls++;
        if (noise) root_noise(t, nd->ps, nd->vs, dir); else normalise(nd->ps);
        nd->ns = 0;
        nd->qsa = (double*)malloc(sizeof(double) * A); nd->nsa = (int64_t*)calloc(A, sizeof(int64_t));
        for (int a = 0; a < A; a++) nd->qsa[a] = NAN_SENTINEL;
        nd->qs = v[0];
        nd->has_ps = 1;
        memcpy(v_out, v, sizeof(float) * (size_t)n);
        return;
    }
    if (noise) root_noise(t, nd->ps, nd->vs, dir);       /* revisited root :150-154 (the noised Ps is stored back :177) */

    int a = pick_highest_ucb(nd, t->args.cpuct, forced, t->step, t->args.fpu);   /* :158-166 */
    int8_t next[7 * 96];
    memcpy(next, s, (size_t)t->sbytes);
    int next_player = spo_make_move(next, &t->rules, a, 0, -1, 0, 0, 0);
    if (next_player != 0) spo_swap_players(next, &t->rules, next_player);

    long idx = nd - t->nodes;                            /* the node array may move during the recursion */
    float v[4], rolled[4];
    search(t, next, 0, 0, NULL, v);                      /* :168 */
    nd = &t->nodes[idx];
    for (int i = 0; i < n; i++) rolled[(i + next_player) % n] = v[i];   /* np.roll(v, next_player) :169 */

    nd->qsa[a] = ((double)nd->nsa[a] * nd->qsa[a] + (double)rolled[0]) / (double)(nd->nsa[a] + 1);   /* :171 */
    nd->qs = ((float)(nd->ns + 1) * nd->qs + rolled[0]) / (float)(nd->ns + 2);                         /* :172, float32 */
    nd->nsa[a] += 1;
    nd->ns += 1;
    memcpy(v_out, rolled, sizeof(float) * (size_t)n);
}

void mo_search(mo_tree* t, const int8_t* canonical, int noise, int forced, int step, const double* dir, float* v_out) {
    t->step = step;
    search(t, canonical, noise, forced, dir, v_out);
}

int mo_get_action_prob(mo_tree* t, const int8_t* canonical, double temp, int full, const double* dir, double* probs, double* q,
                       int64_t* nsa_out, double* qsa_out, long* ns_out, float* qs_out) {
    const int n = t->n;
    int nb = full ? t->args.num_sims : t->args.num_sims / t->args.ratio_full;     /* :55 */
    int forced = full && t->args.forced_playouts;                                  /* :56 */
    float v[4];
    for (int step = 0; step < nb; step++) {
        int noise = (step == 0 && full && t->args.dirichlet_noise);
        if (noise && !dir) return -1;
        t->step = step;
        search(t, canonical, noise, forced, dir, v);
    }
    node* nd = lookup(t, canonical);
    if (!nd || !nd->has_ps) return -2;
    double counts[A];
    for (int a = 0; a < A; a++) counts[a] = (double)nd->nsa[a];
    if (q) for (int p = 0; p < n; p++) q[p] = p == 0 ? (double)nd->qs : (double)(-nd->qs / (float)(n - 1));   /* :65-66, float32 under NumPy 2 scalar rules */
    if (forced) {                                                                   /* :69-74 */
        double best = 0;
        for (int a = 0; a < A; a++) if (counts[a] > best) best = counts[a];
        for (int a = 0; a < A; a++) {
            double c = counts[a];
            if (c != best) c = c - (double)(int64_t)sqrt(KFORCED * (double)nd->ps[a] * (double)nb);
            counts[a] = c > 1 ? c : 0;
        }
    }
    if (probs) {
        if (temp == 0) {                                                            /* :87-92 (first best instead of a random one) */
            double best = -1; int ba = 0;
            for (int a = 0; a < A; a++) if (counts[a] > best) { best = counts[a]; ba = a; }
            for (int a = 0; a < A; a++) probs[a] = a == ba ? 1.0 : 0.0;
        } else {
            double sum = 0;
            for (int a = 0; a < A; a++) { counts[a] = pow(counts[a], 1.0 / temp); sum += counts[a]; }   /* :94-97 */
            for (int a = 0; a < A; a++) probs[a] = counts[a] / sum;
        }
    }
    if (nsa_out) memcpy(nsa_out, nd->nsa, sizeof(int64_t) * A);
    if (qsa_out) memcpy(qsa_out, nd->qsa, sizeof(double) * A);
    if (ns_out) *ns_out = nd->ns;
    if (qs_out) *qs_out = nd->qs;
    return 0;
}
