"""Self-play entry point: drop-in for the reference's self_play.py (play_game :37-82, play :85-119,
`python -m custom_alphazero.self_play` :122-188).

play()  plays ConfigB200.games_per_iteration games as ONE batch on the GPU (the reference fans out
        os.cpu_count()-1 processes, one game each) and returns the same arrays: states float32
        [S, H, W, 4] (parent positions), policies float64 [S, A], rewards int [S].
play_game()  one game through the compat MCTS class, same signature and return values as the
        reference (plumbing / correctness surface; it pays a host round trip per simulation).
"""
import os
import time
from typing import Dict, List, Optional, Tuple

import numpy as np

from az_b200.engine import Rules
from az_b200.selfplay import SelfPlayRunner
from custom_alphazero import paths
from custom_alphazero.config import ConfigB200, ConfigConnectN, ConfigGeneral, ConfigMCTS, ConfigSelfPlay
from custom_alphazero.mcts.mcts import MCTS
from custom_alphazero.serving.factory import append_queue, get_run_id
from custom_alphazero.utils import HostModel, best_saved_model, best_saved_model_hash, reset_plays_inferences_dict

if ConfigGeneral.game == "chess":  # self_play.py:24-27 of the reference
    from custom_alphazero.chess.board import Board
    from custom_alphazero.chess.move import Move
    from custom_alphazero.chess.utils import get_all_possible_moves
elif ConfigGeneral.game == "connect_n":
    from custom_alphazero.connect_n.board import Board
    from custom_alphazero.connect_n.move import Move

    get_all_possible_moves = Board.get_all_possible_moves
else:
    raise NotImplementedError


def _alternating_rewards(final_result, n_plies):
    """self_play.py:69-78 of the reference: the last mover gets the game result (1 win, 0 draw), every other ply
    going backwards gets its negation; then the (disabled by default) discount by distance to the end."""
    z = np.repeat(final_result, n_plies)
    z[-2::-2] = -z[-2::-2]
    return z * ConfigSelfPlay.discounting_factor ** np.arange(n_plies)[::-1]


def play_game(process_id: int, all_possible_moves: List[Move], mcts_iterations: int, run_id: str,
              plays_inferences: Optional[Dict[str, Tuple[np.ndarray, float]]] = None
              ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, MCTS]:
    """One game through the drop-in MCTS class; signature and return values of the reference's play_game
    (self_play.py:37-82): states [T, H, W, 4] float32 (positions BEFORE each move), policies [T, A] float64,
    rewards [T], and the search object with its model detached."""
    np.random.seed(int((process_id + 1) * time.time()) % (2**32 - 1))  # every worker samples differently
    evaluator = None if ConfigGeneral.http_inference else HostModel(best_saved_model(run_id))
    search = MCTS(board=Board(), all_possible_moves=all_possible_moves, concurrency=ConfigGeneral.concurrency,
                  plays_inferences=plays_inferences, model=evaluator, use_solver=ConfigMCTS.use_solver)
    positions, targets = [], []
    while not search.board.is_game_over():
        search.search(mcts_iterations)
        played_greedily = search.board.fullmove_number >= ConfigMCTS.index_move_greedy
        before, _after, target, _move = search.play(played_greedily, return_details=True)
        positions.append(before)
        targets.append(target)
    rewards = _alternating_rewards(search.board.get_result(keep_same_player=True), len(positions))
    search.model = None  # the reference drops it so the object can be pickled across processes
    return np.asarray(positions), np.asarray(targets), rewards, search


_live = {}


def _runner(net, games, game_id_base):
    """A fresh runner per iteration: game ids (the Philox counters) and the weights are baked into the
    captured CUDA graphs, so a new iteration means a new capture (~0.2 s against seconds of self-play)."""
    _live.clear()  # release the previous iteration's slab before allocating the next
    rules = Rules(ConfigConnectN.board_width, ConfigConnectN.board_height, ConfigConnectN.n, bool(ConfigConnectN.gravity))
    r = SelfPlayRunner(rules, n_trees=min(ConfigB200.concurrent_games, games), sims_per_move=ConfigSelfPlay.mcts_iterations,
                       net=net, games_target=games, game_id_base=game_id_base, seed=ConfigB200.seed,
                       move_mode="philox", auto_restart=True, unroll=ConfigB200.graph_unroll,
                       max_free_sims=ConfigB200.max_free_sims, fin_capacity=games,
                       index_move_greedy=ConfigMCTS.index_move_greedy,
                       dirichlet_noise=bool(ConfigMCTS.enable_dirichlet_noise), dirichlet_alpha=ConfigMCTS.dirichlet_noise_value,
                       dirichlet_ratio=ConfigMCTS.dirichlet_noise_ratio)
    _live["runner"] = r
    return r


def play(run_id: str, plays_inferences: Optional[Dict[str, Tuple[np.ndarray, float]]] = None,
         iteration: int = 0) -> Tuple[np.ndarray, np.ndarray, np.ndarray, list]:
    """One self-play iteration on the GPU.  plays_inferences is accepted for signature compatibility:
    the evaluation cache of the reference only saves CPU net calls and never changes results."""
    games = ConfigB200.games_per_iteration
    if ConfigGeneral.game == "chess":
        return _play_chess(run_id, games, iteration)
    runner = _runner(best_saved_model(run_id), games, game_id_base=iteration * games)
    runner.run_until_done()
    states, policies, rewards = runner.collect()
    rewards = rewards * ConfigSelfPlay.discounting_factor ** 0  # discounting_factor == 1 in the reference config
    return states, policies, rewards, []


def _play_chess(run_id: str, games: int, iteration: int):
    """The same iteration for chess: ChessSelfPlayRunner (az_chess_step / az_chess_move around the bf16 net); states
    float32 [S, 8, 8, 118], policies float64 [S, 1880], rewards int [S] of the finished games."""
    from az_b200.chess_selfplay import ChessSelfPlayRunner

    _live.clear()
    r = ChessSelfPlayRunner(n_trees=min(ConfigB200.concurrent_games, games), sims_per_move=ConfigSelfPlay.mcts_iterations,
                            net=best_saved_model(run_id), games_target=games, game_id_base=iteration * games,
                            seed=ConfigB200.seed, move_mode="philox", auto_restart=True, unroll=ConfigB200.graph_unroll,
                            max_free_sims=ConfigB200.max_free_sims, index_move_greedy=ConfigMCTS.index_move_greedy,
                            max_plies=ConfigB200.chess_max_plies)
    _live["runner"] = r
    states, policies, rewards, known = r.run_until_done()
    return states[known], policies[known], rewards[known].astype(np.int64), []


def main(max_iterations: Optional[int] = None):
    plays_inferences = reset_plays_inferences_dict()
    run_id = get_run_id()
    if run_id is None:
        run_id = time.strftime("standalone_%Y%m%d_%H%M%S")
        print(f"No serving process reachable: running stand-alone with id={run_id}")
    print(f"Starting self play with id={run_id}")
    iteration, previous_hash = 0, None
    while max_iterations is None or iteration < max_iterations:
        t0 = time.time()
        os.makedirs(paths.get_self_play_iteration_path(run_id, iteration), exist_ok=True)
        current_hash = best_saved_model_hash(run_id)
        if previous_hash != current_hash:
            plays_inferences, previous_hash = reset_plays_inferences_dict(), current_hash
        states, policies, rewards, _ = play(run_id, plays_inferences, iteration)
        if ConfigSelfPlay.exclude_null_games:
            keep = rewards != 0
            states, policies, rewards = states[keep], policies[keep], rewards[keep]
        print(f"Collected {len(states)} samples in {time.time() - t0:.2f} seconds")
        if (iteration + 1) % ConfigSelfPlay.samples_checkpoint_frequency == 0:
            np.savez(paths.get_self_play_samples_path(run_id, iteration), states=states, policies=policies, values=rewards)
        append_queue(states, policies, rewards)
        iteration += 1


if __name__ == "__main__":
    main(int(os.environ["AZ_MAX_ITERATIONS"]) if "AZ_MAX_ITERATIONS" in os.environ else None)
