// az_bitboard.cuh - Connect-N position as two bitboards (K2) and its NN encoding (K3).
//
// Reference behaviour restated here (paths relative to /root/reference/custom_alphazero/):
//   connect_n/board.py:113-124  legal moves and their order
//   connect_n/board.py:130-146  action list (gravity: x; free placement: x-major x*H + y)
//   connect_n/board.py:178-208  win / draw detection from the last stone
//   connect_n/board.py:210-250  push + play(keep_same_player=True): colours swap after every move
//   connect_n/board.py:83-98    full_state planes
//
// Layout: cell (y, x), row 0 on top, lives at bit y*(W+1) + x.  The extra column per row is a
// sentinel that is never set, so walking a diagonal or a row off the board edge always meets a
// zero bit and no wrap-around test is needed.  H*(W+1) <= 128 bits: one 64-bit word per colour for
// boards up to 6x7 / 7x8, two words for 9x9.  `cur` = stones of the side to move (the reference's
// +1 after mirroring), `opp` = the other side's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace az {

constexpr int kMaxDim = 11;

struct Rules {
    int W, H, n, gravity;
    int A;       // action space (board.py:130-146)
    int cells;   // W*H
    int stride;  // W+1
    int bits;    // H*stride
    uint64_t colmask[kMaxDim][2];  // all cells of column x
    uint64_t boardmask[2];         // all real cells
};

// Compile-time rules for the headline configuration: the same member names as Rules (static members
// are reachable through an instance), so every rule function below is written once, `template <class R>`;
// with FixedRules all divisions by W / stride and all loop bounds fold to constants.
template <int W_, int H_, int N_, int G_>
struct FixedRules {
    static constexpr int W = W_, H = H_, n = N_, gravity = G_;
    static constexpr int A = G_ ? W_ : W_ * H_;
    static constexpr int cells = W_ * H_, stride = W_ + 1, bits = H_ * (W_ + 1);
    static_assert(H_ * (W_ + 1) <= 64, "FixedRules is the one-word fast path");
    __host__ __device__ static constexpr uint64_t col0() {
        uint64_t m = 0;
        for (int y = 0; y < H_; ++y) m |= 1ull << (y * (W_ + 1));
        return m;
    }
    __host__ __device__ static constexpr uint64_t board() {
        uint64_t m = 0;
        for (int x = 0; x < W_; ++x) m |= col0() << x;
        return m;
    }
};

template <class R>
struct is_fixed_rules { static constexpr bool value = false; };
template <int W_, int H_, int N_, int G_>
struct is_fixed_rules<FixedRules<W_, H_, N_, G_>> { static constexpr bool value = true; };

template <int NW>
struct BB {
    uint64_t w[NW];
};

template <int NW>
struct Pos {
    BB<NW> cur, opp;
};

template <int NW>
__host__ __device__ __forceinline__ bool bb_test(const BB<NW>& b, int i) {
    if (NW == 2 && i >= 64) return (b.w[NW - 1] >> (i - 64)) & 1ull;
    return (b.w[0] >> i) & 1ull;
}

template <int NW>
__host__ __device__ __forceinline__ void bb_set(BB<NW>& b, int i) {
    if (NW == 2 && i >= 64)
        b.w[NW - 1] |= 1ull << (i - 64);
    else
        b.w[0] |= 1ull << i;
}

template <int NW>
__host__ __device__ __forceinline__ BB<NW> bb_or(const BB<NW>& a, const BB<NW>& b) {
    BB<NW> r;
#pragma unroll
    for (int i = 0; i < NW; ++i) r.w[i] = a.w[i] | b.w[i];
    return r;
}

#ifdef __CUDACC__
__device__ __forceinline__ int popc64(uint64_t v) { return __popcll(v); }

// index of the j-th (0-based) set bit of a 64-bit word; caller guarantees it exists
__device__ __forceinline__ int nth_set64(uint64_t v, int j) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    int c = __popc(lo);
    if (j < c) return (int)__fns(lo, 0, j + 1);
    return 32 + (int)__fns(hi, 0, j - c + 1);
}

template <int NW>
__device__ __forceinline__ int nth_set(const BB<NW>& b, int j) {
    if (NW == 2) {
        int c = popc64(b.w[0]);
        if (j >= c) return 64 + nth_set64(b.w[NW - 1], j - c);
    }
    return nth_set64(b.w[0], j);
}

// Legal moves in BOARD order (board.py:113-124).
//   gravity: bit x of the result = column x has an empty top cell (ascending x)
//   free:    every empty cell, ascending bit index == row-major (y, x)
template <int NW, class R>
__device__ __forceinline__ BB<NW> legal_set(const R& r, const Pos<NW>& p) {
    BB<NW> occ = bb_or(p.cur, p.opp), l;
    if (r.gravity) {
        l.w[0] = ~occ.w[0] & ((1ull << r.W) - 1ull);
        if (NW == 2) l.w[NW - 1] = 0;
    } else {
        if constexpr (is_fixed_rules<R>::value) {
            l.w[0] = ~occ.w[0] & R::board();
        } else {
#pragma unroll
            for (int i = 0; i < NW; ++i) l.w[i] = ~occ.w[i] & r.boardmask[i];
        }
    }
    return l;
}

template <int NW>
__device__ __forceinline__ int bb_count(const BB<NW>& b) {
    int c = popc64(b.w[0]);
    if (NW == 2) c += popc64(b.w[NW - 1]);
    return c;
}

// j-th legal move in board order -> (cell bit, action index).  kWarpUniform: all 32 lanes of the warp
// call with the same position and j (tree kernels), which lets a ballot find the j-th legal column;
// the per-thread environment kernels pass false.
template <bool kWarpUniform = true, int NW, class R>
__device__ __forceinline__ void edge_move(const R& r, const Pos<NW>& p, const BB<NW>& legal, int j, int& bit,
                                          int& action) {
    if (r.gravity) {
        const uint32_t lm = (uint32_t)legal.w[0];
        int x;
        if (kWarpUniform) {  // lane x votes when column x is legal and has rank j
            const int ln = (int)(threadIdx.x & 31);
            const unsigned vote = __ballot_sync(0xffffffffu, ((lm >> ln) & 1u) && __popc(lm & ((1u << ln) - 1u)) == j);
            x = __ffs((int)vote) - 1;
        } else {  // j-th set bit of at most 11: peel j lowest bits
            uint32_t m = lm;
            for (int i = 0; i < j; ++i) m &= m - 1;
            x = __ffs((int)m) - 1;
        }
        int filled;
        if constexpr (is_fixed_rules<R>::value) {
            filled = popc64((p.cur.w[0] | p.opp.w[0]) & (R::col0() << x));
        } else {
            filled = popc64((p.cur.w[0] | p.opp.w[0]) & r.colmask[x][0]);
            if (NW == 2) filled += popc64((p.cur.w[NW - 1] | p.opp.w[NW - 1]) & r.colmask[x][1]);
        }
        bit = (r.H - 1 - filled) * r.stride + x;  // board.py:212-226: lowest empty row
        action = x;
    } else {
        bit = nth_set(legal, j);
        int y = bit / r.stride, x = bit - y * r.stride;
        action = x * r.H + y;  // board.py:138-146: product(range(W), range(H))
    }
}

// legality of action a in ACTION-LIST order (board.py:154-155)
template <int NW, class R>
__device__ __forceinline__ bool action_legal(const R& r, const Pos<NW>& p, const BB<NW>& legal, int a) {
    if (a >= r.A) return false;
    if (r.gravity) return (legal.w[0] >> a) & 1ull;
    int x = a / r.H, y = a - x * r.H;
    return bb_test(legal, y * r.stride + x);
}

// Number of the mover's stones in line through `bit` along bit-delta d (board.py:186-203).
template <int NW, class R>
__device__ __forceinline__ int run_through(const R& r, const BB<NW>& m, int bit, int d) {
    int run = 1;
    for (int i = bit + d; i < r.bits && run < r.n && bb_test(m, i); i += d) ++run;
    for (int i = bit - d; i >= 0 && run < r.n && bb_test(m, i); i -= d) ++run;
    return run;
}

// Board.play(move, keep_same_player=True) for the stone at `bit` (board.py:233-250):
// returns 0 = game goes on, 1 = the mover connected n (is_null False), 2 = draw (is_null True).
template <int NW, class R>
__device__ __forceinline__ int place(const R& r, Pos<NW>& p, int bit) {
    BB<NW> mover = p.cur;
    bb_set(mover, bit);
    p.cur = p.opp;  // mirror: the opponent becomes the side to move (+1)
    p.opp = mover;
    const int s = r.stride;
    // config.py:47 directions (dx,dy): (0,1) vertical, (1,1), (1,0) horizontal, (1,-1)
    if (NW == 1) {
        // one-word boards: starts of n-runs by shift-and (the sentinel column stops wrap-around), kept
        // only where the run passes through the new stone - same answer as the scan of board.py:186-203
        const uint64_t m = mover.w[0], stone = 1ull << bit;
        const int d4[4] = {s, s + 1, 1, s - 1};
        bool win = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int d = d4[k];
            uint64_t run = m, through = stone;  // through: possible run starts bit, bit-d, ... (negative ones fall off)
            for (int i = 1; i < r.n; ++i) {
                run &= m >> (i * d);
                through |= stone >> (i * d);
            }
            win = win || (run & through) != 0;
        }
        if (win) return 1;
    } else if (run_through(r, mover, bit, s) >= r.n || run_through(r, mover, bit, s + 1) >= r.n ||
               run_through(r, mover, bit, 1) >= r.n || run_through(r, mover, bit, s - 1) >= r.n) {
        return 1;
    }
    BB<NW> l = legal_set(r, p);  // board.py:206-208: no move left -> draw
    bool any = l.w[0] != 0;
    if (NW == 2) any = any || l.w[NW - 1] != 0;
    return any ? 0 : 2;
}

// cell index (row-major, 0..W*H-1) -> plane code: 0 empty, 1 side to move, 2 opponent
template <int NW, class R>
__device__ __forceinline__ int cell_code(const R& r, const Pos<NW>& p, int cell) {
    int y = cell / r.W, x = cell - y * r.W, b = y * r.stride + x;
    return bb_test(p.cur, b) ? 1 : (bb_test(p.opp, b) ? 2 : 0);
}
#endif  // __CUDACC__

inline Rules make_rules(int W, int H, int n, int gravity) {
    Rules r{};
    r.W = W;
    r.H = H;
    r.n = n;
    r.gravity = gravity;
    r.A = gravity ? W : W * H;
    r.cells = W * H;
    r.stride = W + 1;
    r.bits = H * (W + 1);
    for (int x = 0; x < W; ++x)
        for (int y = 0; y < H; ++y) {
            int b = y * r.stride + x;
            r.colmask[x][b >> 6] |= 1ull << (b & 63);
            r.boardmask[b >> 6] |= 1ull << (b & 63);
        }
    return r;
}

}  // namespace az
