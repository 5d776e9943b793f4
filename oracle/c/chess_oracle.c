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
