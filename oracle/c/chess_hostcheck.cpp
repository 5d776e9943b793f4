// chess_hostcheck.cpp - TEST INFRASTRUCTURE ONLY.
// Compiles the device rules header (custom-alphazero_b200/csrc/az_chess.cuh) for the HOST so that the very code
// the CUDA kernels run can be compared with the independent mailbox oracle (c/chess_oracle.c) in the CPU test
// suite, where there is no GPU.  Nothing in the product links this file.
#include "../../custom-alphazero_b200/csrc/az_chess.cuh"

using namespace azc;

extern "C" {

// moves of the side to move as a 1 880-bit mask (30 words); returns the count; flags[0] = in check, flags[1] = unlisted
int hc_legal(const Pos* p, uint64_t* mask_out, int* flags) {
    MoveMask mm;
    int unlisted = 0;
    GenInfo gi = legal_moves(*p, mm, &unlisted);
    for (int i = 0; i < kMaskWords; ++i) mask_out[i] = mm.w[i];
    flags[0] = gi.in_check;
    flags[1] = unlisted;
    return gi.n_moves;
}

void hc_play(const Pos* p, int from, int to, int promo, int keep_same_player, Pos* out) {
    *out = play(*p, from, to, promo, keep_same_player != 0);
}

void hc_mirror(const Pos* p, Pos* out) { *out = mirror(*p); }

int hc_status(const Pos* p) {
    MoveMask mm;
    GenInfo gi = legal_moves(*p, mm, nullptr);
    return game_status(*p, gi);
}

// perft on the keep_same_player path (white to move, mirror after every move)
uint64_t hc_perft_mirrored(const Pos* p, int depth) {
    MoveMask mm;
    GenInfo gi = gen_white(*p, mm);
    if (depth <= 1) return depth == 1 ? (uint64_t)gi.n_moves : 1;
    uint64_t total = 0;
    for (int w = 0; w < kMaskWords; ++w)
        for (u64 b = mm.w[w]; b; b &= b - 1) {
            int mv = act_move(w * 64 + lsb(b));
            Pos q = play(*p, mv & 63, (mv >> 6) & 63, mv >> 12, true);
            total += hc_perft_mirrored(&q, depth - 1);
        }
    return total;
}

int hc_piece_at(const Pos* p, int sq) { return piece_at(*p, sq); }
void hc_start(Pos* out) { *out = start_position(); }
int hc_act_move(int a) { return act_move(a); }
int hc_act_index(int from, int to) { return act_index(from, to); }
}
