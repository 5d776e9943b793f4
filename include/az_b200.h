/* az_b200.h - C ABI of the B200-native self-play search engine (libaz_b200.so).
 *
 * This is the drop-in boundary for the reference's self-play hot path.  The reference
 * (neuronest/custom-alphazero) is pure Python and has no FFI of its own; each entry point below
 * names the reference interface it replaces (paths relative to custom_alphazero/ in the
 * reference).  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, a config struct; no torch / C++ types
 *   - every pointer marked "dev" is DEVICE memory owned by the caller (the Python host allocates
 *     it through torch); the library allocates nothing on the device
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it, never
 *     synchronises and is CUDA-graph capturable
 *   - return value: 0 on success, otherwise an AZ_ERR_* code (host-side argument errors).
 *     Device-side conditions (node pool exhausted, sqrt table exceeded, illegal action) are
 *     sticky per-tree bits in the `status` array that the host reads once per move
 *   - one engine per GPU per process; games are sharded across ranks by game id
 */
#ifndef AZ_B200_H
#define AZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AZ_ABI_VERSION 1
#define AZ_MAX_ACTIONS 128 /* W*H <= 121, bitboard needs H*(W+1) <= 128 */
#define AZ_MAX_DEPTH 128

enum {
    AZ_OK = 0,
    AZ_ERR_ARG = 1,      /* bad configuration / null pointer */
    AZ_ERR_SLAB = 2,     /* slab too small or misaligned */
    AZ_ERR_CUDA = 3,     /* a CUDA runtime call failed (message via az_last_error) */
    AZ_ERR_NO_DEVICE = 4 /* no CUDA device: there is no CPU fallback */
};

/* evaluator that turns a leaf position into (priors, value) */
enum {
    AZ_EVAL_EXTERNAL = 0, /* the policy/value net (or any host evaluator) between two az_step calls */
    AZ_EVAL_UNIFORM = 1,  /* in-kernel: np.full(A, 1/A), value 0 (SURVEY 8c fixed evaluator) */
    AZ_EVAL_HASH = 2      /* in-kernel: oracle/evaluators.py hash evaluator */
};
/* arithmetic of the prior normalisation (mcts/utils.py:4-16) */
enum {
    AZ_PRIOR_F64 = 0, /* float64 evaluator output (serving/factory.py:55) */
    AZ_PRIOR_F32 = 1  /* float32 net output normalised in float32, then widened (mcts.py:131-137) */
};
/* how MCTS.play picks the edge (mcts.py:198-201) */
enum {
    AZ_MOVE_ARGMAX = 0,       /* play(deterministic=True) */
    AZ_MOVE_HOST_UNIFORMS = 1, /* np.random.choice with the draw supplied by the host: uniforms[tree][ply] */
    AZ_MOVE_PHILOX = 2        /* same sampling rule, draw = Philox4x32-10(seed, game_id, ply) on device */
};
/* dtypes of evaluator tensors */
enum { AZ_F32 = 0, AZ_F64 = 1, AZ_BF16 = 2 };

/* per-tree status word (dev int32): phase in the low byte, sticky error bits above */
enum {
    AZ_PHASE_IDLE = 0,    /* no game assigned */
    AZ_PHASE_SEARCH = 1,  /* simulations still to run for the current move */
    AZ_PHASE_READY = 2,   /* sims_per_move reached: waiting for az_play */
    AZ_PHASE_STALLED = 3, /* game finished but the finished-game ring is full */
    AZ_PHASE_MASK = 0xff,
    AZ_FLAG_POOL_OVERFLOW = 1 << 8, /* node pool exhausted: results of this tree are invalid */
    AZ_FLAG_LUT_OVERFLOW = 1 << 9,  /* sum of visits exceeded the pow-half table */
    AZ_FLAG_ILLEGAL = 1 << 10       /* illegal action / inconsistent root */
};

/* Replaces the class-attribute configuration the hot path reads (config.py:19-56):
 * ConfigConnectN.{board_width, board_height, n, gravity}, ConfigSelfPlay.mcts_iterations,
 * ConfigMCTS.{exploration_constant, index_move_greedy}. */
typedef struct az_config {
    int32_t abi_version;       /* AZ_ABI_VERSION */
    int32_t width, height;     /* board; H*(W+1) <= 128 */
    int32_t n_connect;         /* stones in a row to win */
    int32_t gravity;           /* 1: Connect-4 style columns (A = W); 0: free placement (A = W*H) */
    int32_t n_trees;           /* concurrent games (one warp each) on this GPU */
    int32_t node_capacity;     /* nodes per tree per pool half (two halves: re-root compaction ping-pongs) */
    int32_t sims_per_move;     /* ConfigSelfPlay.mcts_iterations */
    int32_t index_move_greedy; /* ConfigMCTS.index_move_greedy (self_play.py:62) */
    int32_t eval_mode;         /* AZ_EVAL_* */
    int32_t prior_mode;        /* AZ_PRIOR_* for the in-kernel evaluators (az_search); az_step follows eval_dtype */
    int32_t move_mode;         /* AZ_MOVE_* */
    int32_t max_free_sims;     /* simulations ending in a terminal leaf that one az_step may run per tree
                                  before handing the batch slot back (they need no evaluation) */
    int32_t fin_capacity;      /* finished-game ring entries */
    int32_t pow_lut_len;       /* entries of the host-built table n -> n ** 0.5 (Q3: libm pow, not sqrt) */
    int32_t auto_restart;      /* 1: a finished game is replaced by the next game id while any remain */
    int32_t inline_play;       /* 1: az_step itself plays the move (the az_play step with the self-play rules) the
                                  moment a tree's budget is spent, whenever the re-root fits in place; az_play
                                  then only serves trees that need the compaction path */
    int32_t eval_cache_log2;   /* 0 = off; else the evaluation memo of the reference (plays_inferences, mcts.py:122-143) as a
                                  direct-mapped table of 2^n entries shared by all trees: a leaf whose position was already
                                  evaluated is expanded on the spot without the evaluator.  Results are identical with or
                                  without it (the evaluator is a pure function of the position); only external evaluations
                                  (az_step / az_advance_fused) are memoised.  az_cache_clear when the weights change. */
    int32_t dirichlet_noise;   /* ConfigMCTS.enable_dirichlet_noise (mcts.py:70-85, off in the reference's config): at the
                                  root every simulation scores the edges with (1 - ratio) * prior + ratio * Dir(alpha * 1_k),
                                  drawn afresh each time (quirk Q4) from Philox(seed, game id, ply, simulation) - the same
                                  distribution as np.random.dirichlet, not the same stream */
    double dirichlet_alpha;    /* ConfigMCTS.dirichlet_noise_value (0.03) */
    double dirichlet_ratio;    /* ConfigMCTS.dirichlet_noise_ratio (0.25) */
    double c_puct;             /* ConfigMCTS.exploration_constant */
    uint64_t seed;             /* Philox key for AZ_MOVE_PHILOX */
    int64_t game_id_base;      /* first global game id of this rank */
    int64_t games_target;      /* games this rank may start in total (auto_restart) */
} az_config;

/* Byte offsets of every array inside the caller-provided slab, so the host can make typed
 * views (torch) of records, finished games and node pools without further calls.
 * Shapes use T = n_trees, C = node_capacity, A = n_actions, P = max_plies = W*H, WD = words. */
typedef struct az_layout {
    size_t total_bytes;
    int32_t n_actions, max_plies, words, max_depth;
    /* per-tree header */
    size_t status;      /* int32  [T] */
    size_t ply;         /* int32  [T]      fullmove_number of the live game board (board.py:40) */
    size_t game_id;     /* int64  [T] */
    size_t root_board;  /* uint64 [T][2][WD]  (side-to-move stones, opponent stones) */
    size_t half;        /* int32  [T]      which pool half holds the live tree */
    size_t root_node;   /* int32  [T]      index of the current root in the live half (re-root is in place
                                           while the half has room for the next search, else the kept subtree
                                           is compacted into the other half and the root is node 0) */
    size_t n_nodes;     /* int32  [T]      nodes used in the live half */
    size_t sims_done;   /* int32  [T] */
    size_t pending;     /* int32  [T]      1 = a leaf awaits its evaluation, 2 = a leaf is selected but not yet handed out */
    size_t path_len;    /* int32  [T] */
    size_t path;        /* int32  [T][max_depth]  node indices root-child ... leaf */
    size_t leaf_board;  /* uint64 [T][2][WD] */
    size_t counters;    /* int64  [T][8]   cumulative: simulations, evaluations, moves, games finished, sum of
                                           selection depths, children created, nodes copied by re-root, memo hits */
    size_t uniforms;    /* double [T][P]   AZ_MOVE_HOST_UNIFORMS draws */
    /* node pools: record A = {double W; int32 N; uint32 link}, link = first_child | k << 24, 0 = no edges */
    size_t node_a;      /* 16 B   [T][2][C] */
    size_t node_p;      /* double [T][2][C]  prior of the edge into the node */
    /* per-tree record of the game in progress (what play_game accumulates, self_play.py:58-66) */
    size_t rec_visits;  /* int32  [T][P][A]  root visit counts by action, -1 = illegal */
    size_t rec_action;  /* int32  [T][P]     action | greedy << 16 */
    size_t rec_board;   /* uint64 [T][P][2][WD]  parent position of each ply */
    size_t rec_len;     /* int32  [T]      plies recorded so far for the game in progress */
    size_t result;      /* int32  [T]      result of the game once it is over (see fin_result) */
    /* finished-game ring */
    size_t fin_count;   /* int32  [1] (+ games_started int64 at fin_count + 8) */
    size_t fin_game_id; /* int64  [F] */
    size_t fin_len;     /* int32  [F] */
    size_t fin_result;  /* int32  [F]   1 = the player who moved last won, 0 = draw (board.py:258-268) */
    size_t fin_visits;  /* int32  [F][P][A] */
    size_t fin_action;  /* int32  [F][P]     action | greedy << 16 */
    size_t fin_board;   /* uint64 [F][P][2][WD] */
    size_t pow_lut;     /* double [pow_lut_len] */
    /* evaluation memo (eval_cache_log2 > 0): S = 2^n entries */
    size_t cache_meta;  /* uint32 [S]   seqlock word: odd = being written, even = version */
    size_t cache_key;   /* uint64 [S][2][WD]  position */
    size_t cache_val;   /* float  [S][A + 1]  priors then value */
} az_layout;

typedef struct az_engine az_engine;

const char *az_last_error(void);
int az_abi_version(void);
/* sizeof(az_config), sizeof(az_layout) as compiled: lets a foreign-language binding verify its mirror */
void az_struct_sizes(size_t *config_bytes, size_t *layout_bytes);

/* Layout / size of the slab for a configuration (host only, no device needed). */
int az_query_layout(const az_config *cfg, az_layout *out);

/* Binds a configuration to a zero-initialised device slab and uploads the pow-half table.
 * Replaces MCTS.__init__ / initialize_root (mcts/mcts.py:89-109) for n_trees trees at once. */
int az_engine_create(const az_config *cfg, void *dev_slab, size_t slab_bytes, const double *host_pow_lut,
                     void *stream, az_engine **out);
void az_engine_destroy(az_engine *e);

/* Starts a fresh game from the empty board in every tree (Board(), board.py:13-42); game ids are
 * game_id_base + tree.  Also clears the finished-game ring and the counters. */
int az_reset_games(az_engine *e, void *stream);
/* Moves the engine's game-id range to [base, base + games_target): the next az_reset_games starts games base + tree.
 * Game ids key the move-sampling / root-noise counters (Philox over seed, game id, ply), so a self-play loop that
 * resets the engine every iteration must move the base or it replays the same games (the reference draws fresh
 * np.random numbers every iteration, self_play.py:37-40).  Host-side only; kernel launches captured in a CUDA graph
 * before the call keep the old base - re-capture. */
int az_set_game_id_base(az_engine *e, int64_t game_id_base);

/* Roots given trees at arbitrary positions: what MCTS(board=...) does with a caller-supplied Board
 * (mcts/mcts.py:98,108-109).  cells: dev int8 [n][H][W] in the reference's convention (+1 = side to
 * move, row 0 = top); plies: dev int32 [n]; tree_ids: dev int32 [n]. */
int az_set_roots(az_engine *e, const int32_t *dev_tree_ids, const int8_t *dev_cells, const int32_t *dev_plies,
                 int32_t n, void *stream);

/* Overrides the per-move simulation budget for subsequent az_step / az_search calls
 * (MCTS.search(iterations_number), mcts/mcts.py:170) and re-arms trees in AZ_PHASE_READY. */
int az_begin_search(az_engine *e, int32_t sims, void *stream);

/* One lock-step advance of every tree with an EXTERNAL evaluator (mcts/mcts.py:170-180):
 *   1. a tree whose leaf was evaluated since the last call expands it with priors/values
 *      (evaluate_and_expand, :145-161) and backs the value up its stored path (backup, :163-168);
 *   2. it then selects the next leaf by PUCT (select, :111-120; UCTEdge terms :39-55); simulations
 *      that end in a terminal leaf are finished on the spot (:179) and, with cfg.inline_play, a spent move
 *      budget is turned into a move right here (the az_play step) - up to max_free_sims such events per call;
 *   3. a non-terminal leaf is encoded as the NN input (Board.full_state, connect_n/board.py:83-98)
 *      into states_out[tree] and leaf_valid_out[tree] = 1.
 * priors: dev [T][A], values: dev [T] (dtype AZ_F32 or AZ_F64); may be NULL on the first call.  As in the
 * reference the dtype decides the arithmetic of normalize_probabilities: AZ_F32 = float32 sum and divide, then
 * widened (the model= path, mcts.py:131-137); AZ_F64 = float64 (the infer_sample path, factory.py:55).
 * states_out: dev [T][H][W][4] (AZ_BF16 or AZ_F32); leaf_valid_out: dev int32 [T]. */
int az_step(az_engine *e, const void *dev_priors, const void *dev_values, int32_t eval_dtype, void *dev_states_out,
            int32_t state_dtype, int32_t *dev_leaf_valid_out, void *stream);
/* az_step that also gathers the leaf batch: leaf_list (dev int32 [n_trees]) receives the indices of the trees whose
 * leaf awaits evaluation (in no particular order), leaf_count (dev int32 [1]) their number.  az_net_forward_gathered
 * evaluates exactly those positions, so a tree without a pending leaf (terminal leaf, game over) costs no net work -
 * the reference evaluates only real leaves too (mcts/mcts.py:122-143). */
int az_step_gather(az_engine *e, const void *dev_priors, const void *dev_values, int32_t eval_dtype, void *dev_states_out,
                   int32_t state_dtype, int32_t *dev_leaf_valid_out, int32_t *dev_leaf_list_out, int32_t *dev_leaf_count_out,
                   void *stream);

/* Simulations that need no evaluator, for trees WITHOUT a leaf in flight (their last simulation ended in a
 * terminal leaf, mcts/mcts.py:179, or spent the move budget): up to max_sims more terminal-leaf simulations /
 * in-line moves per tree, stopping at the first leaf that does need the evaluator, which is parked and submitted
 * by the next az_step / az_advance_fused.  Touches no evaluator buffer, so it may run on a forked stream beside the
 * net; it must complete before the next az_step / az_advance_fused / az_play of the same engine. */
int az_extra_sims(az_engine *e, int32_t max_sims, void *stream);

/* Runs the remaining simulations of the current move for every tree inside ONE kernel using the
 * in-kernel evaluator (eval_mode UNIFORM or HASH): MCTS.search(n) with the fixed evaluator. */
int az_search(az_engine *e, void *stream);

/* MCTS.play (mcts/mcts.py:182-222) for every tree in AZ_PHASE_READY: root policy from visit counts,
 * edge choice, record (parent position, visit counts, action), move applied to the live board
 * (Board.play(..., keep_same_player=True), connect_n/board.py:233-250), re-root to the chosen child
 * keeping its subtree (in place while the live pool half has room for another search, else compacted
 * breadth-first into the other half).  greedy_override: -1 = the self-play rule
 * ply >= index_move_greedy, 0 / 1 = force; move_mode_override: -1 = cfg.move_mode.
 * Finished games are moved to the ring and, with auto_restart, replaced (play_game, self_play.py:59-78). */
int az_play(az_engine *e, int32_t greedy_override, int32_t move_mode_override, void *stream);

/* Forgets every memoised evaluation (the reference resets plays_inferences when the best-model hash changes,
 * self_play.py:142-150). */
int az_cache_clear(az_engine *e, void *stream);

/* Empties the finished-game ring after the host copied it; trees in AZ_PHASE_STALLED hand their game over at
 * the next az_play. */
int az_fin_clear(az_engine *e, void *stream);

/* Standalone environment kernels (K2/K3), batched over n boards in the reference's cell convention.
 * Replaces Board.play(move, keep_same_player=True) (board.py:233-250) incl. push (:210-231) and
 * update_game_over (:178-208); Board.legal_moves_mask (:154-155); Board.full_state (:83-98).
 *   cells_in/out: dev int8 [n][H][W]; actions: dev int32 [n] (index into get_all_possible_moves, :130-146)
 *   status_out: dev int32 [n]: 0 ongoing, 1 mover won, 2 draw, -1 illegal action (board unchanged)
 *   legal_out: dev uint8 [n][A]; states_out: dev float32 [n][H][W][4] */
int az_env_play(const az_config *cfg, const int8_t *dev_cells_in, const int32_t *dev_actions, int32_t n,
                int8_t *dev_cells_out, int32_t *dev_status_out, void *stream);
int az_env_legal(const az_config *cfg, const int8_t *dev_cells, int32_t n, uint8_t *dev_legal_out, void *stream);
int az_env_encode(const az_config *cfg, const int8_t *dev_cells, int32_t n, float *dev_states_out, void *stream);

/* Turns finished-game records (finished-ring layout) into the training arrays play_game returns
 * (self_play.py:66-78), one sample per ply:
 *   states_out   float32 [S][H][W][4]  parent position of the ply (Board.full_state, board.py:83-98)
 *   policies_out float64 [S][A]        N / sum N, or one-hot at the first maximum in edge order when the
 *                                      ply was played greedily; 0 for illegal actions (mcts.py:189-197, 210-214)
 *   values_out   int32   [S]           result * (+1 for the last ply, alternating backwards) (self_play.py:71-78)
 * boards: dev uint64 [G][P][2][WD]; visits: dev int32 [G][P][A]; actions: dev int32 [G][P]; lens, results:
 * dev int32 [G]; offsets: dev int32 [G] = index of the first sample of game g (exclusive prefix sum of lens). */
int az_decode_samples(const az_config *cfg, const uint64_t *dev_boards, const int32_t *dev_visits,
                      const int32_t *dev_actions, const int32_t *dev_lens, const int32_t *dev_results,
                      const int32_t *dev_offsets, int32_t n_games, float *dev_states_out, double *dev_policies_out,
                      int32_t *dev_values_out, void *stream);

/* ---- the two memory-bound ends of the policy/value net (the tower runs through cuDNN) ----
 * Weights are float32 with the BatchNormalization inference transform folded in. */
typedef struct az_head_weights {
    const float *conv_w;   /* dev [3][C]: rows 0-1 policy 1x1 conv, row 2 value 1x1 conv */
    const float *conv_b;   /* dev [3] */
    const float *policy_w; /* dev [A][2*H*W + 1]: Dense(A) on the NHWC-flattened policy planes, rows padded to an odd
                              stride (the kernel copies it verbatim into conflict-free shared memory); 16 B aligned */
    const float *policy_b; /* dev [A] */
    const float *value1_w; /* dev [H*W][256]: Dense(256) weights TRANSPOSED (input cell major); 16 B aligned */
    const float *value1_b; /* dev [256], 16 B aligned */
    const float *value2_w; /* dev [256], 16 B aligned */
    const float *value2_b; /* dev [1] */
} az_head_weights;

/* Stem of ResidualTower (model/tensorflow/model.py:36-46): Conv3x3(4 -> C) + BN + ReLU.
 * states: dev bf16 [n][H][W][4] (what az_step writes); w: dev float [C][4][3][3]; b: dev float [C];
 * out: dev bf16 [n][H][W][C].  C must be 128 (config.py:71). */
int az_net_stem(const void *dev_states, const float *dev_w, const float *dev_b, int32_t n, int32_t height,
                int32_t width, int32_t channels, void *dev_out, void *stream);

/* The same stem on tcgen05 (csrc/az_gemm.cu): states as above; w_bf16: dev bf16 [128][64] with K index = tap * 4 + plane
 * (tap = ky * 3 + kx), zero beyond 36; bias: dev float [128]; H * W <= 128. */
int az_net_stem_tc(const void *dev_states, const void *dev_w_bf16, const float *dev_bias, int32_t n, int32_t height, int32_t width,
                   int32_t channels, void *dev_out, void *stream);

/* PolicyHead + ValueHead (model/tensorflow/model.py:68-149) on the tower output x: dev bf16 [n][H*W][C].
 * priors_out: dev float [n][A] (softmax), values_out: dev float [n] (tanh) - the buffers az_step reads. */
int az_net_heads(const void *dev_x, const az_head_weights *weights, int32_t n, int32_t cells, int32_t channels,
                 int32_t n_actions, float *dev_priors_out, float *dev_values_out, void *stream);

/* The 1x1 projection shortcut of a residual block (model/tensorflow/base_layers.py:105-113, BN folded):
 * y[rows][128] = x[rows][128] . w[128][128]^T, bf16 in / float32 accumulate / bf16 out, no bias (the caller folds it
 * into the bias of the convolution that consumes y).  Hand-written tcgen05 GEMM (csrc/az_gemm.cu).
 * x, y: dev bf16 [rows][channels] (NHWC activations flattened over cells); w: dev bf16 [channels out][channels in];
 * all 16-byte aligned; channels must be 128. */
int az_net_conv1x1(const void *dev_x, const void *dev_w, int64_t rows, int32_t channels, void *dev_y, void *stream);

/* The whole residual tower (model/tensorflow/model.py:48-66: `depth` x ResidualBlock with projection shortcut,
 * base_layers.py:85-125, BN folded) in ONE persistent tcgen05 kernel (csrc/az_tower.cu):
 *   per block  h = relu(conv3x3(x, w1) + b1);  x' = relu(conv3x3(h, w2) + conv1x1(x, wp) + b2p)
 * Activations stay in shared memory / TMEM from the first block to the last; replaces the 12 library convolutions.
 * x, y: dev bf16 [n][H*W][128] (NHWC; y may not alias x).  w_img: dev bf16, depth x 38 stages of 16 KB, each
 * [8 chunks][128 output channels][8 input channels] in consumption order - per block: conv1 = 2 halves of the input
 * channels x taps (ky, kx) row-major, the shortcut x 2 halves, conv2 like conv1 (az_b200/net.py pack_tower_weights).
 * layout 0: one CTA per tile (the stage layout above).  layout 1: CTA pairs (cta_group::2, M = 256 over two tiles): every
 * 16 KB stage is split by output channel, [2 halves][8 chunks][64 output channels][8 input channels], and CTA r of a
 * pair streams half r - per SM half the weight traffic through shared memory (pack_tower_weights(pair=True)).
 * bias: dev float [depth][2][128] (b1; b2 + shortcut bias).  channels must be 128, H * W <= 128,
 * (128 / (H * W)) * W + 1 <= 22 (6x7, 8x8, ...), depth 1..4.  All pointers 16-byte aligned. */
int az_net_tower(const void *dev_x, const void *dev_w_img, const float *dev_bias, int32_t n, int32_t H, int32_t W,
                 int32_t channels, int32_t depth, int32_t layout, void *dev_y, void *stream);

/* Measurement aid: while dev_buffer (dev int64 [148][8], or NULL to switch off) is set, every az_net_tower / az_net_forward
 * launch writes per CTA {cycles of the MMA warp's tile loop, of which waiting for activations, waiting for weight stages,
 * cycles epilogue warp 2 waited for an accumulator, cycles of its epilogue bodies} (tools/time_tower.py). */
int az_net_tower_timing(void *dev_buffer);

/* Measurement aid, the net-side twin of az_debug_timeline: successive az_net_tower / az_net_forward[_gathered] launches
 * record {first CTA start, last CTA end} (globaltimer, ns) into dev_slots[2 * (launch % n_slots)], initialised by the
 * caller to {UINT64_MAX, 0}; NULL switches it off.  The slot is chosen at launch (or graph-capture) time
 * (tools/explore_timeline2.py: where the time of one advance goes inside graph replays). */
int az_net_debug_timeline(void *dev_slots, int32_t n_slots);

/* Head weights of az_net_forward: plain row-major float32, BN folded (no padding or transposition). */
typedef struct az_net_head_params {
    const float *conv_w;   /* dev [3][128]: rows 0-1 policy 1x1 conv (model.py:68-85), row 2 value 1x1 conv (:106-123) */
    const float *conv_b;   /* dev [3] */
    const float *policy_w; /* dev [A][2*H*W]: Dense(A) on the NHWC-flattened policy planes (model.py:86-103) */
    const float *policy_b; /* dev [A] */
    const float *value1_w; /* dev [256][H*W]: Dense(256) (model.py:129-139) */
    const float *value1_b; /* dev [256] */
    const float *value2_w; /* dev [256]: Dense(1) (model.py:140-149) */
    const float *value2_b; /* dev [1] */
} az_net_head_params;

/* The whole policy/value net - PolicyValueModel.call (model/tensorflow/model.py:182-188): stem Conv3x3(4 -> 128) + BN +
 * ReLU, `depth` residual blocks, PolicyHead softmax and ValueHead tanh - in ONE persistent tcgen05 kernel (the net mode
 * of csrc/az_tower.cu): 8 bytes per cell in, A + 1 floats per position out, no activation ever written to HBM.
 * states: dev bf16 [n][H*W][4] (Board.full_state, what az_step writes); w_img: dev bf16 = 3 stem stages (9 taps x
 * [128][16 K], planes in K 0-3) followed by the az_net_tower image (az_b200/net.py pack_stem_weights /
 * pack_tower_weights); stem_bias: dev float [128]; tower_bias: dev float [depth][2][128];
 * priors: dev float [n][A]; values: dev float [n] - the buffers az_step consumes.
 * layout: as az_net_tower (0 = one CTA per tile, 1 = CTA pairs; the stem stages are split the same way).
 * H * W <= 48, (128 / (H*W)) * A <= 32, channels 128, depth 1..4; states 8-byte, w_img 16-byte aligned. */
int az_net_forward(const void *dev_states, const void *dev_w_img, const float *dev_stem_bias, const float *dev_tower_bias,
                   const az_net_head_params *heads, int32_t n, int32_t H, int32_t W, int32_t channels, int32_t depth,
                   int32_t n_actions, int32_t layout, float *dev_priors, float *dev_values, void *stream);
/* az_net_forward on a gathered batch (az_step_gather): position i of the batch is tree dev_index[i], i < *dev_count
 * (read on the device: no host round trip); states / priors / values stay indexed by tree, n_max = their first
 * dimension.  Trees that are not listed keep their old priors / values. */
int az_net_forward_gathered(const void *dev_states, const void *dev_w_img, const float *dev_stem_bias,
                            const float *dev_tower_bias, const az_net_head_params *heads, const int32_t *dev_index,
                            const int32_t *dev_count, int32_t n_max, int32_t H, int32_t W, int32_t channels, int32_t depth,
                            int32_t n_actions, int32_t layout, float *dev_priors, float *dev_values, void *stream);

/* az_net_forward_gathered with the idle trees of `engine` simulating INSIDE the same kernel (two more warps per CTA):
 * every tree without a leaf in flight - its last simulations ended in terminal leaves (mcts/mcts.py:179) and used up
 * az_step's max_free_sims - runs up to max_sims more evaluator-free simulations while the tensor cores evaluate the
 * others' leaves; the first leaf that does need the evaluator is parked and handed out by the next az_step (as
 * az_extra_sims parks it); moves are left to the next az_step.  The same simulations in the same order, only earlier:
 * per-tree results do not change.  The net kernel owns its SMs (227 KB of shared memory per CTA), so no other kernel can
 * run beside it; it issues an instruction in only 15 % of its cycles, which is what these warps use.  The engine must be
 * the plain 6x7 connect-4 configuration (compile-time rules, no root noise), else AZ_ERR_ARG. */
int az_net_forward_trees(const void *dev_states, const void *dev_w_img, const float *dev_stem_bias,
                         const float *dev_tower_bias, const az_net_head_params *heads, const int32_t *dev_index,
                         const int32_t *dev_count, int32_t n_max, int32_t H, int32_t W, int32_t channels, int32_t depth,
                         int32_t n_actions, int32_t layout, float *dev_priors, float *dev_values, az_engine *engine,
                         int32_t max_sims, void *stream);

/* The dense layers of az_net_heads alone, on the output of az_net_head_convs: hd dev float [n][cells][3] -> priors / values
 * as above (same weights struct; conv_w / conv_b unused).  az_net_head_convs + az_net_heads_dense = az_net_heads with the
 * convolutions read in 128-bit loads and accumulated in float32. */
int az_net_heads_dense(const float *dev_hd, const az_head_weights *weights, int32_t n, int32_t cells, int32_t n_actions,
                       float *dev_priors_out, float *dev_values_out, void *stream);

/* Only the two 1x1 head convolutions + BN + ReLU (model.py:76-80, :114-118) of az_net_heads, for shapes whose dense
 * layers do not fit its shared memory (chess: 64 cells x 1 880 actions; they then run through cuBLAS).
 * x: dev bf16 [n][cells][C]; conv_w: dev float [3][C] (rows 0-1 policy, row 2 value); conv_b: dev float [3];
 * out: dev float [n][cells][3]. */
int az_net_head_convs(const void *dev_x, const float *dev_conv_w, const float *dev_conv_b, int32_t n, int32_t cells,
                      int32_t channels, float *dev_out, void *stream);

/* The dense layers of both heads for wide action spaces (chess) in one tcgen05 kernel (csrc/az_gemm.cu): policy
 * Dense(A) + softmax (model/tensorflow/model.py:86-103) and value Dense(256) + ReLU + Dense(1) + tanh (:129-149) on
 * the output of az_net_head_convs.
 *   hd: dev float [n][cells][3]; policy_w: dev bf16 [ceil(A / 128) * 128][2 * cells], rows >= A zero; policy_b: dev float [A];
 *   value1_w: dev bf16 [256][cells]; value1_b: dev float [256]; value2_w: dev float [256]; value2_b: dev float [1];
 *   priors_out: dev float [n][A]; values_out: dev float [n]; scratch: dev float [n][AZ_DENSE_HEAD_SPLITS][2] (partial
 *   softmax statistics between the two launches this call makes).  cells must be 64, A a multiple of 4. */
#define AZ_DENSE_HEAD_SPLITS 4
int az_net_dense_heads(const float *dev_hd, const void *dev_policy_w, const float *dev_policy_b, const void *dev_value1_w,
                       const float *dev_value1_b, const float *dev_value2_w, const float *dev_value2_b, int32_t n,
                       int32_t cells, int32_t n_actions, float *dev_priors_out, float *dev_values_out, float *dev_scratch,
                       void *stream);

/* The three per-tree stages between two passes of the tower in ONE launch (one warp per tree):
 * az_net_heads on the tower output of each tree's pending leaf, az_step with those priors / value (kept in
 * registers), az_net_stem on the newly selected leaf.  Same results as the three calls in sequence.
 *   tower_out: dev bf16 [T][H*W][128] = the tower's output for the leaves handed out by the previous call
 *              (NULL on the first call);  stem_out: dev bf16 [T][H*W][128] = the tower's next input
 *   leaf_valid_out: dev int32 [T]: 1 = stem_out[tree] holds a fresh leaf. */
int az_advance_fused(az_engine *e, const void *dev_tower_out, const az_head_weights *weights, const float *dev_stem_w,
                     const float *dev_stem_b, void *dev_stem_out, int32_t *dev_leaf_valid_out, void *stream);

/* Test aid: n Dirichlet(alpha * 1_k) samples from the device sampler used for the root noise
 * (out: dev double [n][k]); sample i uses the Philox counter (game = i, ply = 0, simulation = 0). */
int az_debug_dirichlet(uint64_t seed, double alpha, int32_t k, int32_t n, double *dev_out, void *stream);

/* Measurement aid: successive az_advance_fused / az_step / az_step_gather launches record {first block start, last warp end} (globaltimer,
 * ns) into dev_slots[2 * (launch % n_slots)], which the caller initialises to {UINT64_MAX, 0}.  NULL switches it
 * off.  The slot is chosen at launch (or graph-capture) time. */
int az_debug_timeline(az_engine *e, void *dev_slots, int32_t n_slots);

/* ---------------------------------------------------------------- chess ----
 * The reference's chess environment (chess/board.py, chess/move.py, chess/utils.py) is a subclass of python-chess's
 * Board; these entry points replace what it uses from it, batched over n boards.  A board is eight 64-bit words
 * (square a1 = bit 0 ... h8 = bit 63, like python-chess): */
#define AZ_CHESS_ACTIONS 1880  /* len(get_all_possible_moves()), chess/utils.py:11-32 */
#define AZ_CHESS_MASK_WORDS 30 /* legal mask over the action list, 64 actions per word */
#define AZ_CHESS_PLANES 118    /* Board.full_state channels, chess/board.py:58-73 */
typedef struct az_chess_pos {
    uint64_t pawns, knights, bishops, rooks, queens, kings; /* both colours */
    uint64_t white;                                         /* squares of white pieces */
    uint64_t meta; /* bits 0-3 castling rights (1 K, 2 Q, 4 k, 8 q); 4-10 en-passant square + 1 (0 none); 11 side to move
                      (0 white); 16-31 halfmove clock; 32-47 fullmove number; 48 repetition flag and 49 valid flag of
                      a history entry (az_chess_encode) */
} az_chess_pos;

/* get_all_possible_moves() (chess/utils.py:11-32) in the reference's order (chess/move.py:28-32):
 * host_moves_out[a] = from | to << 6 | promo << 12, promo 0 none, 1 'b', 2 'n', 3 'q', 4 'r' (host memory, 1880 entries). */
int az_chess_action_table(uint16_t *host_moves_out);

/* Board.moves as Board.legal_moves_mask (chess/board.py:46-48, :116-117) + is_game_over / get_result (:178-190).
 *   mask_out:   dev uint64 [n][30] or NULL: bit a = action a is legal for the side to move
 *   count_out:  dev int32 [n] or NULL: number of legal moves (including black promotions, which the reference's action
 *               list cannot express)
 *   status_out: dev int32 [n] or NULL: bits 0-1: 0 ongoing, 1 checkmate (the side to move lost), 2 draw (insufficient
 *               material, stalemate, 75-move rule); bit 2: side to move is in check; bits 8+: unlisted moves */
int az_chess_legal(const az_chess_pos *dev_pos, int32_t n, uint64_t *dev_mask_out, int32_t *dev_count_out,
                   int32_t *dev_status_out, void *stream);

/* Board.play(move, keep_same_player) (chess/board.py:162-173): push_uci, then mirror() with the turn forced to white.
 * actions: dev int32 [n] indices into the action list; status_out: dev int32 [n]: state of the NEW position as above
 * (0 / 1 / 2), or -1 for an illegal action (board copied unchanged; python-chess raises). */
int az_chess_play(const az_chess_pos *dev_pos_in, const int32_t *dev_actions, int32_t n, int32_t keep_same_player,
                  az_chess_pos *dev_pos_out, int32_t *dev_status_out, void *stream);

/* Board.full_state (chess/board.py:58-73): states_out dev [n][8][8][118] float32 (AZ_F32) or bf16 (AZ_BF16).
 * history: dev az_chess_pos [n][7], the 7 older entries of Board.state_history (oldest first; an entry without the
 * valid flag is the zero padding), or NULL for what the deque holds on the keep_same_player path. */
int az_chess_encode(const az_chess_pos *dev_pos, const az_chess_pos *dev_history, int32_t n, int32_t dtype,
                    void *dev_states_out, void *stream);

/* Known-answer test and throughput measurement of the move generator: number of move paths of length `depth`
 * (<= 8) from each position along the self-play path (move, mirror, move, ...), equal to the published perft counts.
 * nodes_out: dev uint64 [n]. */
int az_chess_perft(const az_chess_pos *dev_pos, int32_t n, int32_t depth, uint64_t *dev_nodes_out, void *stream);

/* ---- chess search engine: the same warp-per-tree PUCT search (mcts/mcts.py:39-222) over chess positions ----
 * One warp owns one game tree.  A node is {double W; int32 N; uint32 link} + double prior + uint16 action (the move
 * INTO the node as an index into the action list); children of a node are contiguous, 8-aligned, in ascending action
 * order, and get their priors by action (the reference pairs p[legal_mask] with python-chess's generation order,
 * which cannot be pinned here - see DESIGN.md).  The position is replayed in registers while descending (move +
 * mirror per level); leaves are classified by the legal-move generator (checkmate: +1 for the player who moved in,
 * any draw: 0; mcts.py:179 with the Connect-N meaning of get_result(keep_same_player=True)).  Selection arithmetic,
 * tie-break, backup, root policy, move sampling and re-root are those of az_step / az_play.
 * Finished plies go to a sample ring (one entry per ply: parent position, legal actions, visit counts, chosen
 * action), finished games to a small ring (game id, length, result); the host joins them by game id to give every
 * sample its value (self_play.py:66-78). */
#define AZ_CHESS_MAX_CHILDREN 224 /* >= 218, the most legal moves a chess position can have */
typedef struct az_chess_config {
    int32_t abi_version;       /* AZ_ABI_VERSION */
    int32_t n_trees;           /* concurrent games (one warp each) */
    int32_t node_capacity;     /* nodes per tree per pool half */
    int32_t sims_per_move;     /* ConfigSelfPlay.mcts_iterations */
    int32_t index_move_greedy; /* ConfigMCTS.index_move_greedy */
    int32_t eval_mode;         /* AZ_EVAL_*: external (az_chess_step) or the in-kernel uniform / hash evaluators */
    int32_t prior_mode;        /* AZ_PRIOR_* of the in-kernel evaluators; az_chess_step follows eval_dtype */
    int32_t move_mode;         /* AZ_MOVE_* */
    int32_t max_free_sims;     /* simulations ending in a terminal leaf one az_chess_step may run per tree */
    int32_t max_plies;         /* a game still running after this many plies is recorded as a draw (the reference has
                                  no cut-off besides the 75-move rule); also the row length of `uniforms` */
    int32_t sample_capacity;   /* entries of the sample ring */
    int32_t fin_capacity;      /* entries of the finished-game ring */
    int32_t pow_lut_len;       /* entries of the host-built table n -> n ** 0.5 */
    int32_t auto_restart;      /* 1: a finished game is replaced by the next game id while any remain */
    double c_puct;             /* ConfigMCTS.exploration_constant */
    uint64_t seed;             /* Philox key for AZ_MOVE_PHILOX */
    int64_t game_id_base;      /* first global game id of this rank */
    int64_t games_target;      /* games this rank may start in total */
} az_chess_config;

/* Byte offsets into the caller-provided slab.  T = n_trees, C = node_capacity, S = sample_capacity, F = fin_capacity,
 * P = max_plies, K = AZ_CHESS_MAX_CHILDREN, D = AZ_MAX_DEPTH. */
typedef struct az_chess_layout {
    size_t total_bytes;
    size_t status;     /* int32 [T]  AZ_PHASE_* | AZ_FLAG_* */
    size_t ply;        /* int32 [T] */
    size_t game_id;    /* int64 [T] */
    size_t root_pos;   /* az_chess_pos [T]  white (the side to move) at the bottom */
    size_t half;       /* int32 [T] */
    size_t root_node;  /* int32 [T] */
    size_t n_nodes;    /* int32 [T] */
    size_t sims_done;  /* int32 [T] */
    size_t pending;    /* int32 [T]  1 = a leaf awaits its evaluation */
    size_t path_len;   /* int32 [T] */
    size_t path;       /* int32 [T][D] */
    size_t leaf_pos;   /* az_chess_pos [T] */
    size_t leaf_mask;  /* uint64 [T][32]  legal mask of the pending leaf (30 words used) */
    size_t counters;   /* int64 [T][8]  simulations, evaluations, moves, games finished, sum of depths, children
                                        created, nodes copied by re-root, pool high-water mark */
    size_t uniforms;   /* double [T][P]  AZ_MOVE_HOST_UNIFORMS draws */
    size_t node_a;     /* 16 B   [T][2][C] */
    size_t node_p;     /* double [T][2][C] */
    size_t node_m;     /* uint16 [T][2][C] */
    size_t smp_count;  /* int32 [1] */
    size_t smp_game;   /* int64 [S] */
    size_t smp_ply;    /* int32 [S] */
    size_t smp_pos;    /* az_chess_pos [S]  parent position of the ply */
    size_t smp_k;      /* int32 [S]  legal moves */
    size_t smp_act;    /* uint16 [S][K]  their action indices, ascending */
    size_t smp_n;      /* int32 [S][K]  root visit counts */
    size_t smp_choice; /* int32 [S]  chosen action | greedy << 16 */
    size_t fin_count;  /* int32 [1] (+ games_started int64 at fin_count + 8) */
    size_t fin_game;   /* int64 [F] */
    size_t fin_len;    /* int32 [F] */
    size_t fin_result; /* int32 [F]  1 = the player who moved last won, 0 = draw */
    size_t pow_lut;    /* double [pow_lut_len] */
} az_chess_layout;

typedef struct az_chess_engine az_chess_engine;
void az_chess_struct_sizes(size_t *config_bytes, size_t *layout_bytes);
int az_chess_query_layout(const az_chess_config *cfg, az_chess_layout *out);
int az_chess_engine_create(const az_chess_config *cfg, void *dev_slab, size_t slab_bytes, const double *host_pow_lut,
                           void *stream, az_chess_engine **out);
void az_chess_engine_destroy(az_chess_engine *e);
/* every tree starts game `game_id_base + tree` from the initial position (MCTS.__init__ / initialize_root) */
int az_chess_reset_games(az_chess_engine *e, void *stream);
/* as az_set_game_id_base */
int az_chess_set_game_id_base(az_chess_engine *e, int64_t game_id_base);
/* trees tree_ids[i] get the root position positions[i] (white to move) and a fresh tree */
int az_chess_set_roots(az_chess_engine *e, const int32_t *dev_tree_ids, const az_chess_pos *dev_positions, int32_t n,
                       void *stream);
/* MCTS.search(sims) is about to run on every live tree: simulation counters to zero, budget = sims (0 keeps it) */
int az_chess_begin_search(az_chess_engine *e, int32_t sims, void *stream);
/* all remaining simulations of the move with the in-kernel evaluator (eval_mode uniform / hash) */
int az_chess_search(az_chess_engine *e, void *stream);
/* one lock-step advance with an external evaluator, like az_step: consume priors dev [T][1880] / values dev [T]
 * (AZ_F32 or AZ_F64; ignored for trees without a pending leaf), simulate up to the next leaf, write its 118 planes
 * to states_out dev bf16 [T][8][8][plane_stride] and leaf_valid_out dev int32 [T].  Only planes [plane_first, 118) are
 * written, at channel index plane - plane_first (plane_first = 0: all of them; 84: the initial-position entry, the
 * current entry and the scalars - on the self-play path the six older history entries are always empty, so nothing is
 * lost and the stem multiplies 34 planes instead of 118); channels beyond 118 - plane_first are written as zeros so a
 * tensor-core stem can read a channel count that is a multiple of 8 in place (plane_stride >= 118 - plane_first). */
int az_chess_step(az_chess_engine *e, const void *dev_priors, const void *dev_values, int32_t eval_dtype,
                  void *dev_states_out, int32_t plane_stride, int32_t plane_first, int32_t *dev_leaf_valid_out, void *stream);
/* states_out may be NULL: then no planes are written and the caller evaluates the leaves from `leaf_pos` in the slab
 * (az_chess_stem below). */

/* Stem of the net (model/tensorflow/model.py:36-46: Conv3x3(118 -> 128) + BN + ReLU) for positions on the self-play
 * path, straight from the 64-byte boards.  There Board.full_state (chess/board.py:58-73) is six empty history entries +
 * the initial position + the current entry, so only 20 planes vary: w_reduced = dev float [128][24][9] holds the folded
 * stem weights of the current entry's 14 planes (98-111), the 6 scalar planes (112-117) and 4 zero planes, [plane][tap]
 * per output channel; cell_map = dev float [2][64][128]: map 0 = folded bias + the initial position's contribution per
 * cell (array order, row 0 = rank 8); map 1 = the folded bias alone, used for the un-mirrored ply-0 root (the start
 * position with a zero halfmove clock), whose deque is seven empty entries + the state (chess/board.py:37-40).
 * out: dev bf16 [n][8][8][128], the tower's input. */
int az_chess_stem(const az_chess_pos *dev_pos, int32_t n, const float *dev_w_reduced, const float *dev_cell_map, void *dev_out,
                  void *stream);
/* The same on tcgen05 (csrc/az_gemm.cu): w_reduced_bf16 = dev bf16 [128][256], K index = tap * 24 + plane (zero beyond
 * 216); two positions per 128-row tile, im2col tile built in shared memory from the boards. */
int az_chess_stem_tc(const void *dev_pos, int32_t n, const void *dev_w_reduced_bf16, const float *dev_cell_map, void *dev_out,
                     void *stream);
/* MCTS.play for every tree whose budget is spent: sample-ring entry, move, re-root (in place, or compacted into the
 * other pool half), game end -> finished ring + next game.  greedy_override / move_mode_override: -1 = configured. */
int az_chess_move(az_chess_engine *e, int32_t greedy_override, int32_t move_mode_override, void *stream);
/* the host has read the rings: both counts back to zero */
int az_chess_rings_clear(az_chess_engine *e, void *stream);
/* sample-ring entries -> training arrays (self_play.py:63-66): states_out dev float32 [n][8][8][118] of the parent
 * positions, policies_out dev float64 [n][1880] = N / sum N, or one-hot at the first maximum for greedy plies. */
int az_chess_decode_samples(const az_chess_pos *dev_pos, const int32_t *dev_k, const uint16_t *dev_act,
                            const int32_t *dev_n, const int32_t *dev_choice, int32_t n, float *dev_states_out,
                            double *dev_policies_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AZ_B200_H */
