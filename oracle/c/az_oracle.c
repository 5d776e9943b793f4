/* az_oracle.c - plain-C restatement of the reference's Connect-N environment, PUCT search and
 * self-play loop.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): linked by tests/ and by
 * bench.py's cpu_baseline leg, never by the product.
 *
 * Citations are to /root/reference/custom_alphazero/<file>:<lines>.
 *
 * It deliberately shares no representation with the CUDA engine: boards are int8 arrays
 * scanned from the last stone (like the reference), the tree is a pointer structure with one
 * heap board per node created eagerly at expansion (like the reference).  Arithmetic that
 * decides parity follows the reference operation by operation in IEEE double: build with
 * -ffp-contract=off -fno-builtin-pow (oracle/Makefile) so that nothing is fused and
 * pow(x, 0.5) stays the libm call CPython's `x ** 0.5` makes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AZO_MAX_ACTIONS 128
#define AZO_C_PUCT 1.5          /* config.py:51 */
#define AZO_INDEX_MOVE_GREEDY 8 /* config.py:55 */

typedef struct {
    int width, height, n, gravity;
} azo_rules;

/* evaluator kinds */
enum { AZO_EVAL_UNIFORM = 0, AZO_EVAL_HASH = 1, AZO_EVAL_CALLBACK = 2 };
/* prior arithmetic: 0 = float64 normalisation (serving/factory.py:55 path),
 *                   1 = float32 normalisation then widening (model= path, mcts.py:131-137,
 *                       under the pinned numpy 1.24 promotion rules) */
enum { AZO_PRIOR_F64 = 0, AZO_PRIOR_F32 = 1 };

/* state is the reference's full_state, float32 [H][W][4] (board.py:83-98) */
typedef void (*azo_eval_cb)(const float *state, double *priors, double *value, void *user);

typedef struct azo_node {
    int8_t *cells;           /* [H*W], row 0 = top; +1 side to move after mirroring */
    int plies, over, drawn;  /* drawn: -1 None, 1 draw, 0 decisive (board.py:41-42) */
    int k;                   /* number of edges; 0 = unexpanded or terminal */
    struct azo_node **child; /* per edge */
    double *prior, *w;
    int *n, *action;         /* action = index into the action list (board.py:130-146) */
} azo_node;

/* ---------------------------------------------------------------- arena */
typedef struct azo_block {
    struct azo_block *next;
    size_t used, cap;
} azo_block;

typedef struct {
    azo_block *head;
} azo_arena;

static void *arena_alloc(azo_arena *a, size_t bytes) {
    bytes = (bytes + 15) & ~(size_t)15;
    if (!a->head || a->head->used + bytes > a->head->cap) {
        size_t cap = bytes > (1u << 20) ? bytes : (1u << 20);
        azo_block *b = (azo_block *)malloc(sizeof(azo_block) + cap);
        b->next = a->head;
        b->used = 0;
        b->cap = cap;
        a->head = b;
    }
    void *p = (char *)(a->head + 1) + a->head->used;
    a->head->used += bytes;
    return p;
}

static void arena_free(azo_arena *a) {
    while (a->head) {
        azo_block *n = a->head->next;
        free(a->head);
        a->head = n;
    }
}

/* ---------------------------------------------------------------- environment */
static int n_actions(const azo_rules *r) { return r->gravity ? r->width : r->width * r->height; }

/* board.py:113-124: legal moves in BOARD order: gravity -> ascending x with empty top cell;
 * otherwise empty cells row-major (y, x).  Writes (x, y) pairs, returns the count. */
static int legal_moves(const azo_rules *r, const int8_t *c, int *xs, int *ys) {
    int k = 0;
    if (r->gravity) {
        for (int x = 0; x < r->width; ++x)
            if (c[x] == 0) { xs[k] = x; ys[k] = -1; ++k; }
    } else {
        for (int y = 0; y < r->height; ++y)
            for (int x = 0; x < r->width; ++x)
                if (c[y * r->width + x] == 0) { xs[k] = x; ys[k] = y; ++k; }
    }
    return k;
}

/* board.py:130-146: index in the action list (gravity: x; else x-major x*H + y) */
static int action_index(const azo_rules *r, int x, int y) { return r->gravity ? x : x * r->height + y; }

/* board.py:178-208 */
static void refresh_over(const azo_rules *r, azo_node *b, int x0, int y0) {
    static const int dirs[4][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}}; /* config.py:47 (dx, dy) */
    if (b->over) return;
    int8_t colour = b->cells[y0 * r->width + x0];
    for (int d = 0; d < 4; ++d) {
        int run = 1;
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
            int dx = dirs[d][0] * sgn, dy = dirs[d][1] * sgn, x = x0, y = y0;
            while (x + dx >= 0 && x + dx < r->width && y + dy >= 0 && y + dy < r->height) {
                if (b->cells[(y + dy) * r->width + (x + dx)] != colour) break;
                ++run;
                x += dx;
                y += dy;
                if (run >= r->n) { b->over = 1; b->drawn = 0; return; }
            }
        }
    }
    int xs[AZO_MAX_ACTIONS], ys[AZO_MAX_ACTIONS];
    if (legal_moves(r, b->cells, xs, ys) == 0) { b->over = 1; b->drawn = 1; }
}

/* board.py:210-250 with keep_same_player=True, in place.  Returns 0 on an illegal move. */
static int play_in_place(const azo_rules *r, azo_node *b, int x, int y) {
    if (b->over) return 1; /* Q7: finished boards are returned unchanged (board.py:239-240) */
    if (r->gravity) {
        y = r->height - 1; /* board.py:212-226: row above the first non-empty cell from the top */
        for (int row = 0; row < r->height; ++row)
            if (b->cells[row * r->width + x] != 0) { y = row - 1; break; }
        if (y < 0) return 0;
    } else if (b->cells[y * r->width + x] != 0) {
        return 0;
    }
    b->cells[y * r->width + x] = 1; /* side to move is always +1 under keep_same_player */
    refresh_over(r, b, x, y);
    b->plies += 1;
    for (int i = 0; i < r->width * r->height; ++i) b->cells[i] = (int8_t)(-b->cells[i]); /* board.py:169-176, 245-246 */
    return 1;
}

static azo_node *new_node(azo_arena *a, const azo_rules *r, const azo_node *from) {
    azo_node *nd = (azo_node *)arena_alloc(a, sizeof(azo_node));
    memset(nd, 0, sizeof(*nd));
    nd->cells = (int8_t *)arena_alloc(a, (size_t)(r->width * r->height));
    if (from) {
        memcpy(nd->cells, from->cells, (size_t)(r->width * r->height));
        nd->plies = from->plies;
        nd->over = from->over;
        nd->drawn = from->drawn;
    } else {
        memset(nd->cells, 0, (size_t)(r->width * r->height));
        nd->drawn = -1;
    }
    return nd;
}

/* board.py:83-98: dstack(eye(3)[cells], ones*turn); eye(3)[-1] is row 2 */
static void full_state(const azo_rules *r, const azo_node *b, float *out) {
    for (int i = 0; i < r->width * r->height; ++i) {
        int8_t v = b->cells[i];
        out[4 * i + 0] = v == 0;
        out[4 * i + 1] = v == 1;
        out[4 * i + 2] = v == -1;
        out[4 * i + 3] = 1.0f;
    }
}

/* ---------------------------------------------------------------- evaluators (oracle/evaluators.py) */
static void eval_hash(const azo_rules *r, const azo_node *b, double *priors, double *value) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (int i = 0; i < r->width * r->height; ++i) {
        int8_t v = b->cells[i];
        uint64_t c = v == 0 ? 0 : (v == 1 ? 1 : 2);
        h = (h ^ (c + 1)) * 0x100000001B3ull;
    }
    int A = n_actions(r);
    for (int a = 0; a < A; ++a) {
        uint64_t m = (h ^ ((uint64_t)a * 0x9E3779B97F4A7C15ull)) * 0xFF51AFD7ED558CCDull;
        priors[a] = (double)(((m >> 40) % 1000) + 1);
    }
    *value = ((double)((h >> 20) % 2001) - 1000.0) / 1000.0;
}

/* ---------------------------------------------------------------- numpy restatements */
/* numpy add.reduce over a contiguous 1-D array = pairwise sum (umath loops_utils):
 * n < 8 left fold; n <= 128 eight strided accumulators combined as a balanced tree, then the tail */
static double np_sum_f64(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double r[8];
    int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

static float np_sum_f32(const float *a, int n) {
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    float r[8];
    int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

/* mcts/utils.py:4-16.  p holds the k legal priors in ACTION-LIST order; result in out. */
void azo_normalise(const double *p, int k, int prior_mode, double *out) {
    if (prior_mode == AZO_PRIOR_F32) {
        float f[AZO_MAX_ACTIONS];
        for (int i = 0; i < k; ++i) f[i] = (float)p[i];
        float s = np_sum_f32(f, k);
        for (int i = 0; i < k; ++i) out[i] = s == 0.0f ? 1.0 / (double)k : (double)(f[i] / s);
    } else {
        double s = np_sum_f64(p, k);
        for (int i = 0; i < k; ++i) out[i] = s == 0.0 ? 1.0 / (double)k : p[i] / s;
    }
}

double azo_pow_half(long long n) { return pow((double)n, 0.5); } /* CPython int ** 0.5 */

/* ---------------------------------------------------------------- search */
typedef struct {
    azo_rules rules;
    azo_arena arena;
    azo_node *board;   /* the live game board (mcts.py:98) */
    azo_node *current; /* current_root (mcts.py:105) */
    int eval_kind, prior_mode;
    azo_eval_cb cb;
    void *user;
    long long sims, evals;
    /* path cache (mcts.py:106) */
    azo_node *path_node[AZO_MAX_ACTIONS + 1];
    int path_edge[AZO_MAX_ACTIONS + 1];
    int path_len;
} azo_search;

/* mcts.py:39-55, evaluated in the reference's order: q + ((c * prior) * total**0.5) / (1 + n) */
static double puct(const azo_node *nd, int e, long long total) {
    double q = nd->n[e] ? nd->w[e] / (double)nd->n[e] : 0.0;
    double u = AZO_C_PUCT * nd->prior[e];
    u = u * pow((double)total, 0.5);
    u = u / (double)(1 + nd->n[e]);
    return q + u;
}

/* mcts.py:64-68: numpy argmax = first maximum */
static int best_edge(const azo_node *nd) {
    long long total = 0;
    for (int e = 0; e < nd->k; ++e) total += nd->n[e]; /* mcts.py:50: sum over parent.edges */
    int best = 0;
    double bv = puct(nd, 0, total);
    for (int e = 1; e < nd->k; ++e) {
        double v = puct(nd, e, total);
        if (v > bv) { bv = v; best = e; }
    }
    return best;
}

/* mcts.py:145-161 */
static double expand(azo_search *s, azo_node *nd) {
    const azo_rules *r = &s->rules;
    int A = n_actions(r);
    double priors[AZO_MAX_ACTIONS], value = 0.0;
    if (s->eval_kind == AZO_EVAL_UNIFORM) {
        for (int a = 0; a < A; ++a) priors[a] = 1.0 / (double)A;
    } else if (s->eval_kind == AZO_EVAL_HASH) {
        eval_hash(r, nd, priors, &value);
    } else {
        float st[AZO_MAX_ACTIONS * 4];
        full_state(r, nd, st);
        s->cb(st, priors, &value, s->user);
    }
    s->evals += 1;
    int xs[AZO_MAX_ACTIONS], ys[AZO_MAX_ACTIONS];
    int k = legal_moves(r, nd->cells, xs, ys);
    /* probabilities[legal_moves_mask]: the legal priors in action-list order (board.py:154-155) */
    uint8_t legal[AZO_MAX_ACTIONS];
    memset(legal, 0, sizeof(legal));
    for (int j = 0; j < k; ++j) legal[action_index(r, xs[j], ys[j])] = 1;
    double sel[AZO_MAX_ACTIONS], norm[AZO_MAX_ACTIONS];
    int m = 0;
    for (int a = 0; a < A; ++a)
        if (legal[a]) sel[m++] = priors[a];
    azo_normalise(sel, k, s->prior_mode, norm);
    nd->k = k;
    nd->child = (azo_node **)arena_alloc(&s->arena, sizeof(azo_node *) * (size_t)k);
    nd->prior = (double *)arena_alloc(&s->arena, sizeof(double) * (size_t)k);
    nd->w = (double *)arena_alloc(&s->arena, sizeof(double) * (size_t)k);
    nd->n = (int *)arena_alloc(&s->arena, sizeof(int) * (size_t)k);
    nd->action = (int *)arena_alloc(&s->arena, sizeof(int) * (size_t)k);
    /* zip(probabilities, board.moves): j-th normalised prior with j-th move in board order (Q1) */
    for (int j = 0; j < k; ++j) {
        azo_node *c = new_node(&s->arena, r, nd);
        play_in_place(r, c, xs[j], ys[j]);
        nd->child[j] = c;
        nd->prior[j] = norm[j];
        nd->w[j] = 0.0;
        nd->n[j] = 0;
        nd->action[j] = action_index(r, xs[j], ys[j]);
    }
    return value;
}

/* mcts.py:170-180 */
static void simulate(azo_search *s) {
    azo_node *nd = s->current;
    s->path_len = 0;
    while (nd->k) { /* mcts.py:111-120 */
        int e = best_edge(nd);
        s->path_node[s->path_len] = nd;
        s->path_edge[s->path_len] = e;
        s->path_len += 1;
        nd = nd->child[e];
    }
    double v;
    if (!nd->over)
        v = -expand(s, nd);
    else
        v = nd->drawn ? 0.0 : 1.0; /* board.py:258-268 with keep_same_player */
    for (int i = s->path_len - 1; i >= 0; --i) { /* mcts.py:163-168 */
        azo_node *p = s->path_node[i];
        int e = s->path_edge[i];
        p->n[e] += 1;
        p->w[e] += v;
        v = -v;
    }
    s->sims += 1;
}

static void search_init(azo_search *s, const azo_rules *r, int eval_kind, int prior_mode, azo_eval_cb cb, void *user) {
    memset(s, 0, sizeof(*s));
    s->rules = *r;
    s->eval_kind = eval_kind;
    s->prior_mode = prior_mode;
    s->cb = cb;
    s->user = user;
    s->board = new_node(&s->arena, r, NULL);
    s->current = new_node(&s->arena, r, s->board);
}

/* ---------------------------------------------------------------- exported entry points */

/* One random playout on the environment only (tests/golden env_* fixtures).
 * idx_fn semantics are in tests/helpers.py (LCG).  cells_out receives the final board. */
int azo_env_playout(const azo_rules *r, uint64_t lcg_state, int *picked, int *n_picked, int *result, int8_t *cells_out) {
    azo_arena a = {0};
    azo_node *b = new_node(&a, r, NULL);
    int xs[AZO_MAX_ACTIONS], ys[AZO_MAX_ACTIONS], np_ = 0;
    while (!b->over) {
        int k = legal_moves(r, b->cells, xs, ys);
        lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
        int idx = (int)((lcg_state >> 33) % (uint64_t)k);
        picked[np_++] = idx;
        if (!play_in_place(r, b, xs[idx], ys[idx])) { arena_free(&a); return -1; }
    }
    *n_picked = np_;
    *result = b->drawn ? 0 : 1;
    memcpy(cells_out, b->cells, (size_t)(r->width * r->height));
    arena_free(&a);
    return 0;
}

/* search(sims) from the position reached by `prefix` actions; root edge statistics out.
 * edge_action/N/W/P must hold n_actions entries; returns the number of root edges. */
int azo_search_once(const azo_rules *r, const int *prefix, int n_prefix, int sims, int eval_kind, int prior_mode,
                    azo_eval_cb cb, void *user, int *edge_action, int *edge_n, double *edge_w, double *edge_p,
                    long long *evals) {
    azo_search s;
    search_init(&s, r, eval_kind, prior_mode, cb, user);
    for (int i = 0; i < n_prefix; ++i) {
        int x = r->gravity ? prefix[i] : prefix[i] / r->height, y = r->gravity ? -1 : prefix[i] % r->height;
        if (!play_in_place(r, s.board, x, y)) { arena_free(&s.arena); return -1; }
    }
    s.current = new_node(&s.arena, r, s.board);
    for (int i = 0; i < sims; ++i) simulate(&s);
    int k = s.current->k;
    for (int e = 0; e < k; ++e) {
        edge_action[e] = s.current->action[e];
        edge_n[e] = s.current->n[e];
        edge_w[e] = s.current->w[e];
        edge_p[e] = s.current->prior[e];
    }
    *evals = s.evals;
    arena_free(&s.arena);
    return k;
}

/* self_play.py:37-82 for one game.
 *   uniforms == NULL -> play(deterministic=True); else one draw per ply (mcts.py:198-201).
 *   visits_out [max_plies][A]: root visit counts scattered by action, -1 for illegal actions.
 *   policy_out [max_plies][A] (may be NULL): the policy target of mcts.py:189-214.
 * Returns the number of plies played (game may be unfinished if max_plies is hit: *result = -2). */
int azo_play_game(const azo_rules *r, int sims, int eval_kind, int prior_mode, azo_eval_cb cb, void *user,
                  const double *uniforms, int max_plies, int *moves_out, int *visits_out, double *policy_out,
                  int *result, long long *sims_done, long long *evals) {
    azo_search s;
    search_init(&s, r, eval_kind, prior_mode, cb, user);
    int A = n_actions(r), ply = 0;
    while (!s.board->over && ply < max_plies) {
        for (int i = 0; i < sims; ++i) simulate(&s);
        azo_node *root = s.current;
        int k = root->k, greedy = s.board->plies >= AZO_INDEX_MOVE_GREEDY; /* self_play.py:62 */
        double pi[AZO_MAX_ACTIONS];
        if (greedy) { /* mcts.py:189-192 */
            int am = 0;
            for (int e = 1; e < k; ++e)
                if (root->n[e] > root->n[am]) am = e;
            for (int e = 0; e < k; ++e) pi[e] = e == am ? 1.0 : 0.0;
        } else { /* mcts.py:193-197 */
            double cnt[AZO_MAX_ACTIONS];
            for (int e = 0; e < k; ++e) cnt[e] = (double)root->n[e];
            azo_normalise(cnt, k, AZO_PRIOR_F64, pi);
        }
        int pick = 0;
        if (!uniforms) { /* mcts.py:198-199 */
            for (int e = 1; e < k; ++e)
                if (pi[e] > pi[pick]) pick = e;
        } else { /* np.random.choice(edges, 1, p=pi): cumsum, /= last, searchsorted side='right' */
            double cdf[AZO_MAX_ACTIONS], acc = 0.0;
            for (int e = 0; e < k; ++e) { acc = e == 0 ? pi[0] : acc + pi[e]; cdf[e] = acc; }
            double last = cdf[k - 1];
            for (int e = 0; e < k; ++e) cdf[e] = cdf[e] / last;
            pick = 0;
            while (pick < k && cdf[pick] <= uniforms[ply]) ++pick;
            if (pick >= k) pick = k - 1; /* unreachable for u < 1 */
        }
        for (int a = 0; a < A; ++a) {
            visits_out[ply * A + a] = -1;
            if (policy_out) policy_out[ply * A + a] = 0.0;
        }
        for (int e = 0; e < k; ++e) {
            visits_out[ply * A + root->action[e]] = root->n[e];
            if (policy_out) policy_out[ply * A + root->action[e]] = pi[e];
        }
        int act = root->action[pick];
        moves_out[ply] = act;
        int x = r->gravity ? act : act / r->height, y = r->gravity ? -1 : act % r->height;
        play_in_place(r, s.board, x, y); /* mcts.py:205 */
        s.current = root->child[pick]; /* mcts.py:207 */
        if (memcmp(s.board->cells, s.current->cells, (size_t)(r->width * r->height)) != 0) { /* mcts.py:208 */
            arena_free(&s.arena);
            return -1;
        }
        ++ply;
    }
    *result = s.board->over ? (s.board->drawn ? 0 : 1) : -2;
    *sims_done = s.sims;
    *evals = s.evals;
    arena_free(&s.arena);
    return ply;
}
