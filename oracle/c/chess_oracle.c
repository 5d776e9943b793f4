/* chess_oracle.c - TEST INFRASTRUCTURE ONLY (never linked or imported by the product).
 *
 * Plain-C restatement of the chess rules the reference gets from python-chess (third party, chess 1.9.4 /
 * python-chess 1.999 in /root/reference/poetry.lock, not vendored, not installed here): call sites
 * /root/reference/custom_alphazero/chess/board.py:46-48 (legal_moves), :166 (push_uci), :168 (mirror),
 * :179-181 (is_game_over / result).  PARITY UNPINNED against python-chess itself; pinned instead on the published
 * perft node counts (tests/test_chess_oracle.py) which any correct move generator must reproduce.
 *
 * Deliberately shares nothing with the CUDA engine's representation: a 64-entry mailbox, pseudo-legal
 * generation per piece with explicit ray walking, legality by playing the move on a copy and asking whether the
 * mover's king is attacked.  Both colours are generated natively (no mirroring).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* piece codes like the reference's Board.array (chess/board.py:100-108): +1..+6 white P N B R Q K, negative black */
typedef struct {
    int8_t sq[64];      /* a1 = 0 ... h8 = 63 */
    int32_t turn;       /* 1 white, 0 black (python-chess: chess.WHITE = True) */
    int32_t castling;   /* 1 K, 2 Q, 4 k, 8 q */
    int32_t ep;         /* en-passant square or -1 */
    int32_t halfmove;
    int32_t fullmove;
} co_state;

typedef struct {
    int8_t from, to, promo; /* promo: 0, or the piece type 2 N, 3 B, 4 R, 5 Q */
    int8_t pad;
} co_move;

static int on_board(int f, int r) { return f >= 0 && f < 8 && r >= 0 && r < 8; }
static int sign(int v) { return (v > 0) - (v < 0); }

static const int KN[8][2] = {{1, 2}, {2, 1}, {-1, 2}, {-2, 1}, {1, -2}, {2, -1}, {-1, -2}, {-2, -1}};
static const int KG[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};

/* is square s attacked by side `by` (+1 white, -1 black)? */
static int attacked(const co_state *b, int s, int by) {
    int f = s & 7, r = s >> 3, i;
    /* pawns: a white pawn attacks one rank up */
    int pr = r - by;
    if (pr >= 0 && pr < 8) {
        if (f > 0 && b->sq[pr * 8 + f - 1] == by * 1) return 1;
        if (f < 7 && b->sq[pr * 8 + f + 1] == by * 1) return 1;
    }
    for (i = 0; i < 8; ++i) {
        int tf = f + KN[i][0], tr = r + KN[i][1];
        if (on_board(tf, tr) && b->sq[tr * 8 + tf] == by * 2) return 1;
        tf = f + KG[i][0];
        tr = r + KG[i][1];
        if (on_board(tf, tr) && b->sq[tr * 8 + tf] == by * 6) return 1;
    }
    for (i = 0; i < 8; ++i) {
        int df = KG[i][0], dr = KG[i][1], tf = f + df, tr = r + dr;
        int diag = df != 0 && dr != 0;
        while (on_board(tf, tr)) {
            int v = b->sq[tr * 8 + tf];
            if (v != 0) {
                if (v == by * 5 || v == by * (diag ? 3 : 4)) return 1;
                break;
            }
            tf += df;
            tr += dr;
        }
    }
    return 0;
}

static int king_square(const co_state *b, int side) {
    int s;
    for (s = 0; s < 64; ++s)
        if (b->sq[s] == side * 6) return s;
    return -1;
}

void co_push(co_state *b, co_move m) {
    int side = b->turn ? 1 : -1;
    int v = b->sq[m.from], cap = b->sq[m.to];
    int old_ep = b->ep;
    b->ep = -1;
    b->halfmove += 1;
    if (abs(v) == 1 || cap != 0) b->halfmove = 0;
    b->sq[m.from] = 0;
    if (abs(v) == 1) {
        if (m.to == old_ep && (m.to & 7) != (m.from & 7) && cap == 0) b->sq[m.to - 8 * side] = 0; /* en passant */
        if (abs(m.to - m.from) == 16) b->ep = (m.from + m.to) / 2;
        if (m.promo) v = side * m.promo;
    }
    if (abs(v) == 6 && abs((m.to & 7) - (m.from & 7)) == 2) { /* castling: hop the rook */
        int r0 = m.from & 56;
        if ((m.to & 7) == 6) {
            b->sq[r0 + 5] = b->sq[r0 + 7];
            b->sq[r0 + 7] = 0;
        } else {
            b->sq[r0 + 3] = b->sq[r0 + 0];
            b->sq[r0 + 0] = 0;
        }
    }
    b->sq[m.to] = (int8_t)v;
    if (m.from == 4 || m.to == 4) b->castling &= ~3;
    if (m.from == 60 || m.to == 60) b->castling &= ~12;
    if (m.from == 7 || m.to == 7) b->castling &= ~1;
    if (m.from == 0 || m.to == 0) b->castling &= ~2;
    if (m.from == 63 || m.to == 63) b->castling &= ~4;
    if (m.from == 56 || m.to == 56) b->castling &= ~8;
    if (!b->turn) b->fullmove += 1;
    b->turn = !b->turn;
}

static int add(co_move *out, int n, int from, int to, int promo) {
    out[n].from = (int8_t)from;
    out[n].to = (int8_t)to;
    out[n].promo = (int8_t)promo;
    out[n].pad = 0;
    return n + 1;
}

static int pseudo(const co_state *b, co_move *out) {
    int side = b->turn ? 1 : -1, n = 0, s, i;
    for (s = 0; s < 64; ++s) {
        int v = b->sq[s], f = s & 7, r = s >> 3, t = abs(v);
        if (v == 0 || sign(v) != side) continue;
        if (t == 1) {
            int tr = r + side, last = side > 0 ? 7 : 0, home = side > 0 ? 1 : 6, df;
            if (tr < 0 || tr > 7) continue;
            if (b->sq[tr * 8 + f] == 0) {
                if (tr == last) {
                    for (i = 5; i >= 2; --i) n = add(out, n, s, tr * 8 + f, i);
                } else {
                    n = add(out, n, s, tr * 8 + f, 0);
                    if (r == home && b->sq[(tr + side) * 8 + f] == 0) n = add(out, n, s, (tr + side) * 8 + f, 0);
                }
            }
            for (df = -1; df <= 1; df += 2) {
                int tf = f + df, to;
                if (tf < 0 || tf > 7) continue;
                to = tr * 8 + tf;
                if (b->sq[to] != 0 && sign(b->sq[to]) == -side) {
                    if (tr == last) {
                        for (i = 5; i >= 2; --i) n = add(out, n, s, to, i);
                    } else {
                        n = add(out, n, s, to, 0);
                    }
                } else if (to == b->ep && b->sq[to] == 0 && r == (side > 0 ? 4 : 3) && b->sq[to - 8 * side] == -side) {
                    n = add(out, n, s, to, 0);
                }
            }
        } else if (t == 2 || t == 6) {
            const int(*d)[2] = t == 2 ? KN : KG;
            for (i = 0; i < 8; ++i) {
                int tf = f + d[i][0], tr = r + d[i][1];
                if (on_board(tf, tr) && sign(b->sq[tr * 8 + tf]) != side) n = add(out, n, s, tr * 8 + tf, 0);
            }
            if (t == 6) {
                int home = side > 0 ? 4 : 60, ks = side > 0 ? 1 : 4, qs = side > 0 ? 2 : 8;
                if (s == home && !attacked(b, s, -side)) {
                    if ((b->castling & ks) && b->sq[s + 3] == side * 4 && b->sq[s + 1] == 0 && b->sq[s + 2] == 0 &&
                        !attacked(b, s + 1, -side) && !attacked(b, s + 2, -side))
                        n = add(out, n, s, s + 2, 0);
                    if ((b->castling & qs) && b->sq[s - 4] == side * 4 && b->sq[s - 1] == 0 && b->sq[s - 2] == 0 &&
                        b->sq[s - 3] == 0 && !attacked(b, s - 1, -side) && !attacked(b, s - 2, -side))
                        n = add(out, n, s, s - 2, 0);
                }
            }
        } else {
            for (i = 0; i < 8; ++i) {
                int df = KG[i][0], dr = KG[i][1], diag = df != 0 && dr != 0, tf, tr;
                if ((t == 3 && !diag) || (t == 4 && diag)) continue;
                tf = f + df;
                tr = r + dr;
                while (on_board(tf, tr)) {
                    int w = b->sq[tr * 8 + tf];
                    if (sign(w) == side) break;
                    n = add(out, n, s, tr * 8 + tf, 0);
                    if (w != 0) break;
                    tf += df;
                    tr += dr;
                }
            }
        }
    }
    return n;
}

/* legal moves of the side to move; returns the count (out must hold 256) */
int co_legal(const co_state *b, co_move *out) {
    co_move tmp[320];
    int side = b->turn ? 1 : -1, n = pseudo(b, tmp), k = 0, i;
    for (i = 0; i < n; ++i) {
        co_state c = *b;
        int ks;
        co_push(&c, tmp[i]);
        ks = king_square(&c, side);
        if (ks < 0 || !attacked(&c, ks, -side)) out[k++] = tmp[i];
    }
    return k;
}

int co_in_check(const co_state *b) {
    int side = b->turn ? 1 : -1, ks = king_square(b, side);
    return ks >= 0 && attacked(b, ks, -side);
}

/* python-chess Board.mirror(): flip vertically, swap colours, castling rights and en-passant square, flip turn */
void co_mirror(co_state *b) {
    co_state c = *b;
    int s;
    for (s = 0; s < 64; ++s) b->sq[s ^ 56] = (int8_t)-c.sq[s];
    b->castling = ((c.castling & 3) << 2) | ((c.castling >> 2) & 3);
    b->ep = c.ep < 0 ? -1 : (c.ep ^ 56);
    b->turn = !c.turn;
}

static int side_insufficient(const co_state *b, int side) {
    int n_side = 0, knights = 0, bishops = 0, majors = 0, s;
    int other_not_kq = 0, any_pawn = 0, any_knight = 0, dark = 0, light = 0;
    for (s = 0; s < 64; ++s) {
        int v = b->sq[s], t = abs(v);
        if (!v) continue;
        if (t == 1) any_pawn = 1;
        if (t == 2) any_knight = 1;
        if (t == 3) {
            if (((s >> 3) + (s & 7)) & 1) light = 1; else dark = 1;
        }
        if (sign(v) == side) {
            ++n_side;
            if (t == 1 || t == 4 || t == 5) majors = 1;
            if (t == 2) knights = 1;
            if (t == 3) bishops = 1;
        } else if (t != 6 && t != 5) {
            other_not_kq = 1;
        }
    }
    if (majors) return 0;
    if (knights) return n_side <= 2 && !other_not_kq;
    if (bishops) return !(dark && light) && !any_pawn && !any_knight;
    return 1;
}

/* 0 ongoing, 1 checkmate (side to move lost), 2 draw: python-chess outcome(claim_draw=False) without repetition */
int co_status(const co_state *b) {
    co_move mv[256];
    int n = co_legal(b, mv);
    if (n == 0 && co_in_check(b)) return 1;
    if (side_insufficient(b, 1) && side_insufficient(b, -1)) return 2;
    if (n == 0) return 2;
    if (b->halfmove >= 150) return 2;
    return 0;
}

uint64_t co_perft(const co_state *b, int depth) {
    co_move mv[256];
    int n = co_legal(b, mv), i;
    uint64_t total = 0;
    if (depth <= 1) return depth == 1 ? (uint64_t)n : 1;
    for (i = 0; i < n; ++i) {
        co_state c = *b;
        co_push(&c, mv[i]);
        total += co_perft(&c, depth - 1);
    }
    return total;
}

/* perft along the reference's self-play path: every move is followed by mirror() (keep_same_player) */
uint64_t co_perft_mirrored(const co_state *b, int depth) {
    co_move mv[256];
    int n = co_legal(b, mv), i;
    uint64_t total = 0;
    if (depth <= 1) return depth == 1 ? (uint64_t)n : 1;
    for (i = 0; i < n; ++i) {
        co_state c = *b;
        co_push(&c, mv[i]);
        co_mirror(&c);
        c.turn = 1;
        total += co_perft_mirrored(&c, depth - 1);
    }
    return total;
}

/* ================================================================== MCTS over chess (mcts/mcts.py:39-222)
 * Pointer tree, one heap node per edge/position, states materialised when a node is first reached (the reference
 * copies a board per child eagerly; the numbers are the same).  Children in ascending action order with priors
 * paired by action - see DESIGN.md on why the reference's pairing for chess cannot be pinned.  Terminal leaves:
 * +1 for the player who moved in on checkmate, 0 for any draw (mcts.py:179 with the Connect-N meaning of
 * get_result(keep_same_player=True); the reference's chess Board does not accept that argument). */
#include <math.h>

typedef struct co_node {
    struct co_node *child; /* k children, contiguous, NULL until expanded */
    co_state *st;          /* position with the side to move as white, NULL until first reached */
    double prior, w;
    int n, k, action;
} co_node;

typedef struct co_chunk {
    struct co_chunk *next;
    size_t used, cap;
} co_chunk;

typedef struct {
    co_chunk *head;
} co_arena;

static void *arena_get(co_arena *a, size_t bytes) {
    bytes = (bytes + 15) & ~(size_t)15;
    if (!a->head || a->head->used + bytes > a->head->cap) {
        size_t cap = bytes > (1u << 22) ? bytes : (1u << 22);
        co_chunk *c = (co_chunk *)malloc(sizeof(co_chunk) + cap);
        c->next = a->head;
        c->used = 0;
        c->cap = cap;
        a->head = c;
    }
    void *p = (char *)(a->head + 1) + a->head->used;
    a->head->used += bytes;
    return p;
}
static void arena_free(co_arena *a) {
    while (a->head) {
        co_chunk *n = a->head->next;
        free(a->head);
        a->head = n;
    }
}

typedef struct {
    int32_t eval_kind;  /* 0 uniform, 1 hash */
    int32_t prior_mode; /* 0 float64, 1 float32 */
    int32_t sims, greedy_idx, max_plies;
    double c_puct;
} co_mcts_cfg;

static const int PROMO_RANK[6] = {0, 0, 2, 1, 4, 3}; /* piece type -> rank of its UCI letter: b < n < q < r */

static double cs_sum_f64(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n > 128) { /* numpy pairwise: split, first half a multiple of 8 */
        int n2 = n / 2;
        n2 -= n2 % 8;
        return cs_sum_f64(a, n2) + cs_sum_f64(a + n2, n - n2);
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
static float cs_sum_f32(const float *a, int n) {
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n > 128) {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return cs_sum_f32(a, n2) + cs_sum_f32(a + n2, n - n2);
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

/* the hash evaluator: FNV-1a over the 64 squares' (piece code + 7), then castling | (ep + 1) << 4 */
static uint64_t hash_state(const co_state *s) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (int sq = 0; sq < 64; ++sq) h = (h ^ (uint64_t)(s->sq[sq] + 7)) * 0x100000001B3ull;
    h = (h ^ (uint64_t)((s->castling & 15) | ((s->ep + 1) << 4))) * 0x100000001B3ull;
    return h;
}

static int cmp_action(const void *a, const void *b) { return ((const int *)a)[0] - ((const int *)b)[0]; }

/* mcts.py:145-161: returns the value of the leaf for the side to move */
static double co_expand(co_arena *ar, const co_mcts_cfg *cfg, co_node *nd, const int16_t *act_index, long long *evals) {
    co_move mv[256];
    int k = co_legal(nd->st, mv);
    int keyed[256][2];
    for (int i = 0; i < k; ++i) {
        keyed[i][0] = act_index[mv[i].from * 64 + mv[i].to] + PROMO_RANK[mv[i].promo];
        keyed[i][1] = i;
    }
    qsort(keyed, (size_t)k, sizeof(keyed[0]), cmp_action);
    double sel[256], value = 0.0;
    if (cfg->eval_kind == 1) {
        uint64_t h = hash_state(nd->st);
        for (int i = 0; i < k; ++i) {
            uint64_t m = (h ^ ((uint64_t)keyed[i][0] * 0x9E3779B97F4A7C15ull)) * 0xFF51AFD7ED558CCDull;
            sel[i] = (double)(((m >> 40) % 1000) + 1);
        }
        value = ((double)((h >> 20) % 2001) - 1000.0) / 1000.0;
    } else {
        for (int i = 0; i < k; ++i) sel[i] = 1.0 / 1880.0;
    }
    *evals += 1;
    double norm[256];
    if (cfg->prior_mode == 1) {
        float f[256];
        for (int i = 0; i < k; ++i) f[i] = (float)sel[i];
        float s = cs_sum_f32(f, k);
        for (int i = 0; i < k; ++i) norm[i] = s == 0.0f ? 1.0 / (double)k : (double)(f[i] / s);
    } else {
        double s = cs_sum_f64(sel, k);
        for (int i = 0; i < k; ++i) norm[i] = s == 0.0 ? 1.0 / (double)k : sel[i] / s;
    }
    nd->child = (co_node *)arena_get(ar, sizeof(co_node) * (size_t)k);
    for (int i = 0; i < k; ++i) {
        co_node *c = nd->child + i;
        c->child = NULL;
        c->st = NULL;
        c->prior = norm[i];
        c->w = 0.0;
        c->n = 0;
        c->k = 0;
        c->action = keyed[i][0];
        /* remember the move in the node's state slot lazily: stored as from/to/promo packed into k while unexpanded */
        c->k = -(1 + (mv[keyed[i][1]].from | (mv[keyed[i][1]].to << 6) | (mv[keyed[i][1]].promo << 12)));
    }
    nd->k = k;
    return value;
}

/* the position of a child, materialised on first use: parent position + move + mirror (chess/board.py:162-173) */
static void co_reach(co_arena *ar, const co_node *parent, co_node *c) {
    if (c->st) return;
    int code = -c->k - 1;
    co_move m;
    m.from = (int8_t)(code & 63);
    m.to = (int8_t)((code >> 6) & 63);
    m.promo = (int8_t)(code >> 12);
    m.pad = 0;
    c->st = (co_state *)arena_get(ar, sizeof(co_state));
    *c->st = *parent->st;
    co_push(c->st, m);
    co_mirror(c->st);
    c->st->turn = 1;
    c->k = 0;
}

static int co_best_edge(const co_mcts_cfg *cfg, const co_node *nd) {
    long long total = 0;
    for (int e = 0; e < nd->k; ++e) total += nd->child[e].n; /* mcts.py:50 */
    int best = 0;
    double bv = 0.0;
    for (int e = 0; e < nd->k; ++e) {
        const co_node *c = nd->child + e;
        double q = c->n ? c->w / (double)c->n : 0.0;
        double u = cfg->c_puct * c->prior;
        u = u * pow((double)total, 0.5);
        u = u / (double)(1 + c->n);
        double v = q + u;
        if (e == 0 || v > bv) {
            bv = v;
            best = e;
        }
    }
    return best;
}

/* One self-play game from `start` (white to move).  Per ply p < returned length: out_k[p] legal moves, their actions
 * out_act[p][0..k) (ascending) and root visit counts out_n[p][0..k), out_choice[p] = action | greedy << 16.
 * uniforms: one draw per ply for np.random.choice, or NULL for play(deterministic=True).
 * out_result: 1 the player who moved last won, 0 draw (also when max_plies cuts the game).  counters: sims, evals. */
int co_mcts_game(const co_mcts_cfg *cfg, const co_state *start, const int16_t *act_index, const double *uniforms,
                 int32_t *out_k, uint16_t *out_act, int32_t *out_n, int32_t *out_choice, int32_t *out_result,
                 long long *counters) {
    co_arena ar = {NULL};
    co_node *root = (co_node *)arena_get(&ar, sizeof(co_node));
    memset(root, 0, sizeof(*root));
    root->st = (co_state *)arena_get(&ar, sizeof(co_state));
    *root->st = *start;
    int ply = 0;
    long long sims = 0, evals = 0;
    *out_result = 0;
    while (ply < cfg->max_plies) {
        if (co_status(root->st) != 0) break;
        for (int it = 0; it < cfg->sims; ++it) {
            co_node *path[512];
            int depth = 0;
            co_node *nd = root;
            while (nd->k > 0 && depth < 512) { /* select (mcts.py:111-120) */
                co_node *c = nd->child + co_best_edge(cfg, nd);
                co_reach(&ar, nd, c);
                path[depth++] = c;
                nd = c;
            }
            int st = co_status(nd->st);
            double v;
            if (st == 0) v = -co_expand(&ar, cfg, nd, act_index, &evals); /* mcts.py:175 */
            else v = st == 1 ? 1.0 : 0.0;                                 /* mcts.py:179 */
            for (int i = depth - 1; i >= 0; --i) { /* backup (mcts.py:163-168) */
                path[i]->n += 1;
                path[i]->w += v;
                v = -v;
            }
            ++sims;
        }
        /* play (mcts.py:182-222) */
        int k = root->k, am = 0;
        for (int j = 1; j < k; ++j)
            if (root->child[j].n > root->child[am].n) am = j;
        int greedy = ply >= cfg->greedy_idx, pick = am;
        if (uniforms && !greedy) {
            double total = 0.0, last = 0.0, acc = 0.0;
            for (int j = 0; j < k; ++j) total += (double)root->child[j].n;
            for (int j = 0; j < k; ++j) {
                double pj = total == 0.0 ? 1.0 / (double)k : (double)root->child[j].n / total;
                last = j == 0 ? pj : last + pj;
            }
            pick = k - 1;
            for (int j = 0; j < k; ++j) {
                double pj = total == 0.0 ? 1.0 / (double)k : (double)root->child[j].n / total;
                acc = j == 0 ? pj : acc + pj;
                if (acc / last > uniforms[ply]) {
                    pick = j;
                    break;
                }
            }
        }
        out_k[ply] = k;
        for (int j = 0; j < 224; ++j) {
            out_act[(size_t)ply * 224 + j] = j < k ? (uint16_t)root->child[j].action : (uint16_t)0xffff;
            out_n[(size_t)ply * 224 + j] = j < k ? root->child[j].n : 0;
        }
        out_choice[ply] = root->child[pick].action | (greedy ? 1 << 16 : 0);
        co_node *next = root->child + pick;
        co_reach(&ar, root, next);
        root = next;
        ++ply;
        int st = co_status(root->st);
        if (st != 0) {
            *out_result = st == 1 ? 1 : 0;
            break;
        }
    }
    counters[0] = sims;
    counters[1] = evals;
    arena_free(&ar);
    return ply;
}
