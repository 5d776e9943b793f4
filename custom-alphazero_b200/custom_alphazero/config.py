"""Hyper-parameters the self-play hot path reads, under the reference's names.

Same class / attribute names and default values as the reference's config.py (ConfigGeneral :7-16,
ConfigSelfPlay :19-23, ConfigChess :26-35, ConfigConnectN :38-47, ConfigMCTS :50-56, ConfigModel :59-71, ConfigServing
:74-95, ConfigPath :98-124), so code written against `custom_alphazero.config` keeps working; only the
entries the hot path or its entry point touch are kept, plus the B200 knobs at the bottom.  As in the
reference these are plain class attributes that may be patched before use.
"""


class ConfigGeneral:
    game = "connect_n"  # self_play / mcts drive connect_n; chess has its own environment and engine (custom_alphazero.chess, az_b200.chess)
    mono_process = False  # kept for API compatibility; games are a GPU batch, not processes
    concurrency = False
    http_inference = False
    self_play_gpu_index = "0"  # reference: "-1" (CPU); here self-play IS the GPU path
    serving_gpu_index = "-1"
    training_gpu_index = "0"


class ConfigSelfPlay:
    mcts_iterations = 250  # simulations per move
    discounting_factor = 1  # 1 disables discounting (self_play.py:75-78)
    exclude_null_games = True  # drawn games contribute no samples (self_play.py:155-162)
    samples_checkpoint_frequency = 1


class ConfigChess:
    piece_symbols = [None, "p", "n", "b", "r", "q", "k"]
    initial_board_fen = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR"
    initial_turn, initial_castling_rights, initial_ep_quare = "w", "KQkq", "-"
    initial_halfmove_clock, initial_fullmove_number = "0", "1"
    board_size, number_unique_pieces = 8, 12


class ConfigConnectN:
    board_width, board_height = 7, 6
    n = 4  # stones in a row to win
    gravity = True
    white, empty, black = 1, 0, -1
    pieces = {1: "X", 0: ".", -1: "O"}
    directions = [(0, 1), (1, 1), (1, 0), (1, -1)]  # (dx, dy) scanned by the win test


class ConfigMCTS:
    exploration_constant = 1.5
    index_move_greedy = 8  # plies from which the move is the most visited one
    enable_dirichlet_noise = False  # off in the reference too; on the GPU the noise comes from a device Philox stream
    dirichlet_noise_value, dirichlet_noise_ratio = 0.03, 0.25
    use_solver = False


class ConfigModel:
    filters, depth = 128, 4
    batch_size, training_epochs = 256, 1
    l2_penalization_term = 1e-4
    momentum = 0.9
    maximum_learning_rate, minimum_learning_rate = 1e-2, 1e-4
    learning_rates = {range(0, 150000): 1e-2, range(150000, 300000): 1e-3}


class ConfigServing:
    serving_host, serving_port = "localhost", 5555
    serving_address = "http://{0}:{1}".format(serving_host, serving_port)
    minimum_training_size, samples_queue_size = 2500, 10000
    inference_batch_size, inference_timeout = 1, 1


class ConfigPath:
    run_id_path = "/api/run-id"
    append_queue_path = "/api/queue/append"
    results_dir, self_play_dir, training_dir, evaluation_dir = "results", "self_play", "training", "evaluation"
    samples_file = "samples.npz"
    model_prefix, model_meta, model_success = "model", "meta.json", "MODEL_SAVED_SUCCESSFULLY"


class ConfigB200:
    """Knobs of the batched GPU engine (no counterpart in the reference)."""

    concurrent_games = 4096  # trees resident on one GPU (one warp each)
    games_per_iteration = 4096  # games one call of self_play.play() finishes (reference: cpu_count - 1)
    graph_unroll = 8  # lock-step advances captured per CUDA graph
    max_free_sims = None  # terminal-leaf simulations one az_step may finish per tree; None = the runner's choice (2 where idle
    # trees go on simulating inside the net kernel, az_net_forward_trees; else 8)
    seed = 0  # Philox key of the move sampler
    eval_cache_log2 = 20  # device memo of leaf evaluations with 2^n entries (the reference's plays_inferences); 0 = off
    chess_max_plies = 512  # chess: a game still running after this many plies is recorded as a draw
