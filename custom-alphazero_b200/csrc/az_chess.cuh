// az_chess.cuh - chess rules as bitboards (BASELINE config C5).
//
// The reference wraps python-chess (/root/reference/custom_alphazero/chess/board.py:12, third party, pinned
// chess 1.9.4 / python-chess 1.999, not vendored and not installed): move generation, push, mirror, game end
// and castling / en-passant bookkeeping all live there.  This header restates the published rules of chess
// and python-chess's documented conventions the reference relies on:
//   squares          a1 = 0 ... h8 = 63, rank-major (chess.SQUARES)
//   Board.moves      chess/board.py:46-48   legal moves of the side to move
//   Board.play       chess/board.py:162-173 push_uci, then with keep_same_player: mirror() (vertical flip, colours,
//                                           castling rights and en-passant square swapped, turn back to white)
//   get_result       chess/board.py:178-190 checkmate -> the side to move lost; every other end is a draw
//   is_game_over     python-chess outcome(claim_draw=False): checkmate, insufficient material, stalemate,
//                    75-move rule (halfmove clock >= 150); fivefold repetition needs the move stack, which
//                    mirror() drops, so it can never fire on the keep_same_player path and is not modelled
//   full_state       chess/board.py:58-73   118 planes
//   action list      chess/utils.py:11-32   1 880 moves, see tools/gen_chess_tables.py
//
// B200 notes: everything is register arithmetic on 64-bit words - sliding attacks by the "o - 2s" subtraction
// trick with BREV for the negative ray (one instruction on the GPU), knight / king / pawn attacks by shifts; the
// only tables are the 8 KB square-pair -> action index and its 3.7 KB inverse.  The generator is fully legal
// (check mask, pin lines, king danger squares with the king lifted), so there is no make/unmake in the tree
// kernels.  Moves always come out as a 1 880-bit legal mask in action order: that mask is the reference's
// `legal_moves_mask` and child j of a tree node is its j-th set bit.
//
// The header compiles for the host as well (oracle/c/chess_hostcheck.cpp links it into a test-only checker so the
// exact device logic is exercised on CPU against the independent mailbox oracle); the product only uses it
// from CUDA kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AZC_HD __host__ __device__ __forceinline__
#define AZC_HDM __host__ __device__ __forceinline__ /* member functions */
#define AZC_TABLE static __device__ const
#else
#define AZC_HD static inline
#define AZC_HDM inline
#define AZC_TABLE static const
#endif

namespace azc {

typedef uint64_t u64;

#if defined(__CUDACC__)
namespace dev_tables {
#include "az_chess_tables.inc"
}
#undef AZC_TABLE
#define AZC_TABLE static const
#endif
namespace host_tables {
#undef AZC_N_ACTIONS
#include "az_chess_tables.inc"
}

constexpr int kActions = 1880;
constexpr int kMaskWords = 30;  // ceil(1880 / 64)
constexpr int kPlanes = 118;

AZC_HD int act_index(int from, int to) {
#if defined(__CUDA_ARCH__)
    return dev_tables::ACT_INDEX[from * 64 + to];
#else
    return host_tables::ACT_INDEX[from * 64 + to];
#endif
}
AZC_HD int act_move(int action) {
#if defined(__CUDA_ARCH__)
    return dev_tables::ACT_MOVE[action];
#else
    return host_tables::ACT_MOVE[action];
#endif
}

// ---- bit helpers ----
AZC_HD int popc(u64 x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
AZC_HD int lsb(u64 x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
AZC_HD u64 brev(u64 x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
    return __builtin_bswap64(x);
#endif
}
// vertical flip: rank r <-> rank 7 - r
AZC_HD u64 vflip(u64 x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((u64)__byte_perm(lo, 0, 0x0123) << 32) | (u64)__byte_perm(hi, 0, 0x0123);
#else
    return __builtin_bswap64(x);
#endif
}

constexpr u64 FILE_A = 0x0101010101010101ull, FILE_B = FILE_A << 1, FILE_G = FILE_A << 6, FILE_H = FILE_A << 7;
constexpr u64 RANK_1 = 0xffull, RANK_3 = RANK_1 << 16, RANK_8 = RANK_1 << 56;
constexpr u64 DARK = 0xaa55aa55aa55aa55ull, LIGHT = ~DARK;
constexpr u64 DIAG = 0x8040201008040201ull, ANTI = 0x0102040810204080ull;

AZC_HD u64 bit(int sq) { return 1ull << sq; }
AZC_HD u64 rank_line(int sq) { return RANK_1 << (sq & 56); }
AZC_HD u64 file_line(int sq) { return FILE_A << (sq & 7); }
AZC_HD u64 diag_line(int sq) {
    int d = (sq >> 3) - (sq & 7);
    return d >= 0 ? DIAG << (8 * d) : DIAG >> (8 * -d);
}
AZC_HD u64 anti_line(int sq) {
    int d = (sq >> 3) + (sq & 7) - 7;
    return d >= 0 ? ANTI << (8 * d) : ANTI >> (8 * -d);
}
// attacks of a slider on `sq` along `line` (which contains sq) with blockers `occ`
AZC_HD u64 line_attacks(u64 occ, int sq, u64 line) {
    u64 s = bit(sq), m = line ^ s, o = occ & m;
    u64 fwd = o - 2 * s;
    u64 rev = brev(brev(o) - 2 * brev(s));
    return (fwd ^ rev) & m;
}
AZC_HD u64 rook_attacks(u64 occ, int sq) { return line_attacks(occ, sq, rank_line(sq)) | line_attacks(occ, sq, file_line(sq)); }
AZC_HD u64 bishop_attacks(u64 occ, int sq) { return line_attacks(occ, sq, diag_line(sq)) | line_attacks(occ, sq, anti_line(sq)); }
AZC_HD u64 knight_attacks(u64 b) {
    u64 l1 = (b >> 1) & ~FILE_H, l2 = (b >> 2) & ~(FILE_G | FILE_H);
    u64 r1 = (b << 1) & ~FILE_A, r2 = (b << 2) & ~(FILE_A | FILE_B);
    u64 h1 = l1 | r1, h2 = l2 | r2;
    return (h1 << 16) | (h1 >> 16) | (h2 << 8) | (h2 >> 8);
}
AZC_HD u64 king_attacks(u64 b) {
    u64 a = ((b << 1) & ~FILE_A) | ((b >> 1) & ~FILE_H);
    b |= a;
    return a | (b << 8) | (b >> 8);
}

// ---- position ----
// meta: bits 0-3 castling rights (1 white king side, 2 white queen side, 4 black king side, 8 black queen side),
//       bits 4-10 en-passant square + 1 (0 = none), bit 11 side to move (0 white, 1 black),
//       bits 16-31 halfmove clock, bits 32-47 fullmove number, bit 48 repetition flag (history entries of the
//       encoder), bit 49 entry is valid (history entries).
struct Pos {
    u64 pawns, knights, bishops, rooks, queens, kings, white, meta;
};
constexpr u64 META_CASTLE = 0xfull, META_TURN = 1ull << 11, META_REP = 1ull << 48, META_VALID = 1ull << 49;

AZC_HD u64 occupied(const Pos& p) { return p.pawns | p.knights | p.bishops | p.rooks | p.queens | p.kings; }
AZC_HD int ep_square(const Pos& p) { return (int)((p.meta >> 4) & 127) - 1; }
AZC_HD int halfmove(const Pos& p) { return (int)((p.meta >> 16) & 0xffff); }
AZC_HD int fullmove(const Pos& p) { return (int)((p.meta >> 32) & 0xffff); }
AZC_HD bool black_to_move(const Pos& p) { return (p.meta & META_TURN) != 0; }

AZC_HD Pos start_position() {
    Pos p;
    p.pawns = 0x00ff00000000ff00ull;
    p.knights = 0x4200000000000042ull;
    p.bishops = 0x2400000000000024ull;
    p.rooks = 0x8100000000000081ull;
    p.queens = 0x0800000000000008ull;
    p.kings = 0x1000000000000010ull;
    p.white = 0xffffull;
    p.meta = 0xfull | (1ull << 32);
    return p;
}

// A root straight out of the constructor, before any play(): the reference's deque then holds seven empty entries and
// the state (chess/board.py:37-40); only boards that went through play() + mirror() carry the initial position in entry
// 6 (oracle/chess_ref.py).  On the self-play path that is exactly the start position with a zero halfmove clock: any
// later recurrence of the start position has a non-zero clock (pawn moves and captures are irreversible).
AZC_HD bool fresh_root(const Pos& p) {
    const Pos s = start_position();
    return p.pawns == s.pawns && p.knights == s.knights && p.bishops == s.bishops && p.rooks == s.rooks && p.queens == s.queens &&
           p.kings == s.kings && p.white == s.white && halfmove(p) == 0 && !black_to_move(p);
}

// python-chess Board.mirror(): vertical flip, colours swapped, castling rights and en-passant square follow,
// side to move flips; the clocks are kept (the mirrored board is a stack-less copy).
AZC_HD Pos mirror(const Pos& p) {
    Pos q;
    u64 occ = occupied(p);
    q.pawns = vflip(p.pawns);
    q.knights = vflip(p.knights);
    q.bishops = vflip(p.bishops);
    q.rooks = vflip(p.rooks);
    q.queens = vflip(p.queens);
    q.kings = vflip(p.kings);
    q.white = vflip(occ & ~p.white);
    u64 m = p.meta;
    u64 c = m & META_CASTLE;
    u64 ep = (m >> 4) & 127;
    if (ep) ep = (((ep - 1) ^ 56) + 1);
    q.meta = (m & ~(META_CASTLE | (127ull << 4))) | ((c >> 2) | ((c & 3) << 2)) | (ep << 4);
    q.meta ^= META_TURN;
    return q;
}

// squares of `by_white ? white : black` pieces attacking `sq` given blockers `occ`
AZC_HD u64 attackers_to(const Pos& p, int sq, u64 occ, bool by_white) {
    u64 side = by_white ? p.white : ~p.white;
    u64 s = bit(sq);
    u64 pawn_src = by_white ? (((s >> 7) & ~FILE_A) | ((s >> 9) & ~FILE_H))   // white pawns one rank below
                            : (((s << 7) & ~FILE_H) | ((s << 9) & ~FILE_A));  // black pawns one rank above
    u64 a = pawn_src & p.pawns;
    a |= knight_attacks(s) & p.knights;
    a |= king_attacks(s) & p.kings;
    a |= bishop_attacks(occ, sq) & (p.bishops | p.queens);
    a |= rook_attacks(occ, sq) & (p.rooks | p.queens);
    return a & side & occ;
}

struct MoveMask {
    u64 w[kMaskWords];
};
AZC_HD void mask_clear(MoveMask& m) {
    for (int i = 0; i < kMaskWords; ++i) m.w[i] = 0;
}
AZC_HD int mask_count(const MoveMask& m) {
    int c = 0;
    for (int i = 0; i < kMaskWords; ++i) c += popc(m.w[i]);
    return c;
}
AZC_HD void emit(MoveMask& m, int from, int to, int promo) {
    int a = act_index(from, to) + promo;
    m.w[a >> 6] |= 1ull << (a & 63);
}

// Where generated moves go.  MaskSink: one thread fills a private MoveMask (thread-per-board kernels, host build).
// The tree kernels use a warp-collective sink instead (az_chess_tree.cuh: the lanes share out the targets of a piece).
struct MaskSink {
    MoveMask& m;
    AZC_HDM void clear() { mask_clear(m); }
    AZC_HDM void one(int from, int to, int promo) { emit(m, from, to, promo); }
    AZC_HDM void targets(int from, u64 t) {
        while (t) {
            int to = lsb(t);
            t &= t - 1;
            emit(m, from, to, 0);
        }
    }
    // pawn targets: a move onto the eighth rank is four promotions
    AZC_HDM void pawn_targets(int from, u64 t) {
        while (t) {
            int to = lsb(t);
            t &= t - 1;
            if (to >= 56) {
                for (int pr = 1; pr <= 4; ++pr) emit(m, from, to, pr);
            } else {
                emit(m, from, to, 0);
            }
        }
    }
    AZC_HDM int count() { return mask_count(m); }
};

struct GenInfo {
    int n_moves;
    bool in_check;
};

// Legal moves of WHITE in `p` (the keep_same_player path only ever has white to move; positions with black to
// move are mirrored by the caller, see legal_moves below).  Returns the number of moves and whether the king is
// in check.  A position without a white king (the reference builds such boards for its action list,
// chess/utils.py:14-29) generates pseudo-legal moves, like python-chess does.
template <class Sink>
AZC_HD GenInfo gen_white_to(const Pos& p, Sink& out) {
    out.clear();
    GenInfo gi;
    gi.n_moves = 0;
    gi.in_check = false;
    const u64 occ = occupied(p), us = p.white & occ, them = occ & ~p.white;
    const u64 kbb = p.kings & us;
    const bool has_king = kbb != 0;
    const int ksq = has_king ? lsb(kbb) : 0;
    const u64 diag_them = (p.bishops | p.queens) & them, orth_them = (p.rooks | p.queens) & them;

    u64 check_mask = ~0ull;  // squares a non-king move must land on
    u64 pinned = 0;
    u64 danger = 0;  // squares the king may not step on
    if (has_king) {
        const u64 occ_nk = occ ^ kbb;
        u64 bp = p.pawns & them;
        danger = ((bp >> 7) & ~FILE_A) | ((bp >> 9) & ~FILE_H);
        danger |= knight_attacks(p.knights & them);
        danger |= king_attacks(p.kings & them);
        for (u64 b = diag_them; b; b &= b - 1) danger |= bishop_attacks(occ_nk, lsb(b));
        for (u64 b = orth_them; b; b &= b - 1) danger |= rook_attacks(occ_nk, lsb(b));

        const u64 k_orth = rook_attacks(occ, ksq), k_diag = bishop_attacks(occ, ksq);
        u64 checkers = (((kbb << 7) & ~FILE_H) | ((kbb << 9) & ~FILE_A)) & bp;
        checkers |= knight_attacks(kbb) & p.knights & them;
        checkers |= k_orth & orth_them;
        checkers |= k_diag & diag_them;
        gi.in_check = checkers != 0;
        if (checkers) {
            if (checkers & (checkers - 1)) {
                check_mask = 0;  // double check: king moves only
            } else {
                int c = lsb(checkers);
                check_mask = checkers;
                if (checkers & k_orth & orth_them)
                    check_mask |= k_orth & rook_attacks(occ, c);
                else if (checkers & k_diag & diag_them)
                    check_mask |= k_diag & bishop_attacks(occ, c);
            }
        }
        // pins: enemy sliders that would see the king through exactly one of our pieces
        u64 snipers = (rook_attacks(them, ksq) & orth_them) | (bishop_attacks(them, ksq) & diag_them);
        for (u64 b = snipers; b; b &= b - 1) {
            int s = lsb(b);
            bool orth = (rank_line(ksq) | file_line(ksq)) & bit(s);
            u64 between = orth ? (rook_attacks(bit(s), ksq) & rook_attacks(kbb, s))
                               : (bishop_attacks(bit(s), ksq) & bishop_attacks(kbb, s));
            u64 blockers = between & occ;
            if (blockers && !(blockers & (blockers - 1)) && (blockers & us)) pinned |= blockers;
        }
    }
    // line through the king and a pinned piece: the only squares that piece may move to
    auto pin_line = [&](int from) -> u64 {
        if (!(pinned & bit(from))) return ~0ull;
        u64 f = bit(from);
        if (rank_line(ksq) & f) return rank_line(ksq);
        if (file_line(ksq) & f) return file_line(ksq);
        if (diag_line(ksq) & f) return diag_line(ksq);
        return anti_line(ksq);
    };

    // king
    if (has_king) {
        out.targets(ksq, king_attacks(kbb) & ~us & ~danger);
        if (!gi.in_check && ksq == 4) {
            // castling (standard chess): rights, empty squares between king and rook, king path not attacked
            if ((p.meta & 1) && (p.rooks & us & bit(7)) && !(occ & 0x60ull) && !(danger & 0x70ull)) out.one(4, 6, 0);
            if ((p.meta & 2) && (p.rooks & us & bit(0)) && !(occ & 0x0eull) && !(danger & 0x1cull)) out.one(4, 2, 0);
        }
    }
    if (check_mask) {
        const u64 tmask = ~us & check_mask;
        for (u64 b = p.knights & us & ~pinned; b; b &= b - 1) {
            int from = lsb(b);
            out.targets(from, knight_attacks(bit(from)) & tmask);
        }
        for (u64 b = (p.bishops | p.queens) & us; b; b &= b - 1) {
            int from = lsb(b);
            out.targets(from, bishop_attacks(occ, from) & tmask & pin_line(from));
        }
        for (u64 b = (p.rooks | p.queens) & us; b; b &= b - 1) {
            int from = lsb(b);
            out.targets(from, rook_attacks(occ, from) & tmask & pin_line(from));
        }
        // pawns
        for (u64 b = p.pawns & us; b; b &= b - 1) {
            int from = lsb(b);
            u64 f = bit(from), pl = pin_line(from);
            u64 t = (f << 8) & ~occ;
            t |= ((t & RANK_3) << 8) & ~occ;
            t |= (((f << 7) & ~FILE_H) | ((f << 9) & ~FILE_A)) & them;
            t &= check_mask & pl;
            out.pawn_targets(from, t);
        }
    }
    // en passant: rare, so legality is checked by playing it on the occupancy
    int ep = ep_square(p);
    if (ep >= 0 && !(occ & bit(ep))) {
        u64 e = bit(ep);
        u64 capturers = (((e >> 7) & ~FILE_A) | ((e >> 9) & ~FILE_H)) & p.pawns & us & (RANK_1 << 32);
        for (u64 b = capturers; b; b &= b - 1) {
            int from = lsb(b);
            bool ok = true;
            if (has_king) {
                u64 cap = bit(ep - 8);
                u64 occ2 = (occ ^ bit(from) ^ cap) | e;
                Pos q = p;
                q.pawns &= ~cap;
                ok = attackers_to(q, ksq, occ2, false) == 0;
            }
            if (ok) out.one(from, ep, 0);
        }
    }
    gi.n_moves = out.count();
    return gi;
}

AZC_HD GenInfo gen_white(const Pos& p, MoveMask& out) {
    MaskSink sink{out};
    return gen_white_to(p, sink);
}

// Plays a move of WHITE given as from / to / promo (promo 0 none, 1 bishop, 2 knight, 3 queen, 4 rook - the
// sorted order of the UCI letters).  Follows python-chess push(): halfmove clock, castling rights,
// en-passant square after a double step, rook hop when castling, captured pawn when taking en passant; the
// fullmove number only advances after a black move (so never on the keep_same_player path).
AZC_HD Pos make_white(const Pos& p, int from, int to, int promo) {
    Pos q = p;
    const u64 f = bit(from), t = bit(to), occ = occupied(p);
    const bool is_pawn = (p.pawns & f) != 0;
    const bool capture = (occ & t) != 0;
    int ep = ep_square(p);
    u64 hm = (u64)halfmove(p) + 1;
    if (is_pawn || capture) hm = 0;
    // remove whatever stands on the target square
    q.pawns &= ~t; q.knights &= ~t; q.bishops &= ~t; q.rooks &= ~t; q.queens &= ~t; q.kings &= ~t;
    q.white &= ~t;
    u64 new_ep = 0;
    if (is_pawn) {
        q.pawns &= ~f;
        if (to == ep && ((to - from) == 7 || (to - from) == 9) && !capture) {
            u64 cap = bit(to - 8);
            q.pawns &= ~cap;
        }
        if (to - from == 16) new_ep = (u64)(from + 8) + 1;
        if (promo == 0) q.pawns |= t;
        else if (promo == 1) q.bishops |= t;
        else if (promo == 2) q.knights |= t;
        else if (promo == 3) q.queens |= t;
        else q.rooks |= t;
    } else if (p.knights & f) {
        q.knights = (q.knights & ~f) | t;
    } else if (p.bishops & f) {
        q.bishops = (q.bishops & ~f) | t;
    } else if (p.rooks & f) {
        q.rooks = (q.rooks & ~f) | t;
    } else if (p.queens & f) {
        q.queens = (q.queens & ~f) | t;
    } else {
        q.kings = (q.kings & ~f) | t;
        if (from == 4 && to == 6 && (p.meta & 1)) {
            q.rooks = (q.rooks & ~bit(7)) | bit(5);
            q.white = (q.white & ~bit(7)) | bit(5);
        } else if (from == 4 && to == 2 && (p.meta & 2)) {
            q.rooks = (q.rooks & ~bit(0)) | bit(3);
            q.white = (q.white & ~bit(0)) | bit(3);
        }
    }
    q.white = (q.white & ~f) | t;
    u64 c = p.meta & META_CASTLE;
    if (from == 4) c &= ~3ull;
    if (from == 7 || to == 7) c &= ~1ull;
    if (from == 0 || to == 0) c &= ~2ull;
    if (to == 63) c &= ~4ull;
    if (to == 56) c &= ~8ull;
    u64 m = p.meta & ~(META_CASTLE | (127ull << 4) | (0xffffull << 16));
    q.meta = (m | c | (new_ep << 4) | (hm << 16)) ^ META_TURN;
    return q;
}

// ---- side-agnostic wrappers (Board.moves / Board.play for either side) ----
// Moves of the side to move as a mask over the action list.  For black the position is mirrored, the moves are
// generated for white and mapped back (ranks flipped); black promotions have no entry in the reference's action
// list (it only holds white ones, chess/utils.py:22-29) - they are reported through `unlisted`.
AZC_HD GenInfo legal_moves(const Pos& p, MoveMask& out, int* unlisted) {
    if (unlisted) *unlisted = 0;
    if (!black_to_move(p)) return gen_white(p, out);
    MoveMask mm;
    GenInfo gi = gen_white(mirror(p), mm);
    mask_clear(out);
    for (int w = 0; w < kMaskWords; ++w)
        for (u64 b = mm.w[w]; b; b &= b - 1) {
            int mv = act_move(w * 64 + lsb(b));
            int from = (mv & 63) ^ 56, to = ((mv >> 6) & 63) ^ 56, pr = mv >> 12;
            if (pr) {
                if (unlisted) ++*unlisted;
            } else {
                emit(out, from, to, 0);
            }
        }
    return gi;
}

// push + optional mirror (chess/board.py:162-173).  The move is given in the coordinates of `p`.
AZC_HD Pos play(const Pos& p, int from, int to, int promo, bool keep_same_player) {
    Pos q;
    if (!black_to_move(p)) {
        q = make_white(p, from, to, promo);  // now black to move
    } else {
        Pos m = make_white(mirror(p), from ^ 56, to ^ 56, promo);
        m.meta += 1ull << 32;  // a black move completes a full move
        q = mirror(m);         // back to the board's own orientation: white to move
    }
    if (keep_same_player) {
        q = mirror(q);
        q.meta &= ~META_TURN;  // "virtually, it is always white to play" (chess/board.py:169)
    }
    return q;
}

// python-chess has_insufficient_material(color) / is_insufficient_material()
AZC_HD bool side_insufficient(const Pos& p, u64 side, u64 other) {
    if (side & (p.pawns | p.rooks | p.queens)) return false;
    if (side & p.knights) return popc(side) <= 2 && !(other & ~p.kings & ~p.queens);
    if (side & p.bishops) {
        bool same = !(p.bishops & DARK) || !(p.bishops & LIGHT);
        return same && !p.pawns && !p.knights;
    }
    return true;
}
AZC_HD bool insufficient_material(const Pos& p) {
    u64 occ = occupied(p), w = occ & p.white, b = occ & ~p.white;
    return side_insufficient(p, w, b) && side_insufficient(p, b, w);
}

// Game state of a position whose legal moves are known: 0 ongoing, 1 checkmate (the side to move lost),
// 2 draw (insufficient material, stalemate, 75-move rule).
AZC_HD int game_status(const Pos& p, const GenInfo& gi) {
    if (gi.n_moves == 0 && gi.in_check) return 1;
    if (insufficient_material(p)) return 2;
    if (gi.n_moves == 0) return 2;
    if (halfmove(p) >= 150) return 2;
    return 0;
}

// Board.array value (chess/board.py:100-108) of a square: 0 empty, +1..+6 white P N B R Q K, -1..-6 black
AZC_HD int piece_at(const Pos& p, int sq) {
    u64 s = bit(sq);
    int v = 0;
    if (p.pawns & s) v = 1;
    else if (p.knights & s) v = 2;
    else if (p.bishops & s) v = 3;
    else if (p.rooks & s) v = 4;
    else if (p.queens & s) v = 5;
    else if (p.kings & s) v = 6;
    return (p.white & s) ? v : -v;
}
// plane of Board.state (chess/board.py:50-56): np.eye(13)[array] wraps negative values, so black pieces land on
// planes 13 - |v| (king 7 ... pawn 12)
AZC_HD int piece_plane(int v) { return v >= 0 ? v : 13 + v; }

}  // namespace azc
