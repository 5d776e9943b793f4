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


def _dist():
    """(rank, world size) of a torchrun launch; (0, 1) for the plain `python -m custom_alphazero.self_play`."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _init_distributed():
    """Under torchrun (RANK / WORLD_SIZE / LOCAL_RANK in the environment): one process per GPU, NCCL (gloo when
    AZ_DIST_BACKEND says so - the CPU tests).  Replaces the joblib fan-out over os.cpu_count()-1 worker processes
    (self_play.py:98-110 of the reference): games are sharded across ranks by id, every rank plays its shard as one
    batch, rank 0 collects."""
    import torch
    import torch.distributed as dist

    if int(os.environ.get("WORLD_SIZE", "1")) <= 1 or dist.is_initialized():
        return
    backend = os.environ.get("AZ_DIST_BACKEND", "nccl")
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    else:
        dist.init_process_group(backend)


def _runner(net, games, game_id_base):
    """A fresh runner per iteration: game ids (the Philox counters) and the weights are baked into the
    captured CUDA graphs, so a new iteration means a new capture (~0.2 s against seconds of self-play).
    The evaluation memo is on, as in the reference (mcts.py:122-143 memoises every evaluation in plays_inferences);
    a fresh runner starts with an empty memo, which is also what the reference does when the weights change
    (self_play.py:142-150)."""
    _live.clear()  # release the previous iteration's slab before allocating the next
    rules = Rules(ConfigConnectN.board_width, ConfigConnectN.board_height, ConfigConnectN.n, bool(ConfigConnectN.gravity))
    r = SelfPlayRunner(rules, n_trees=max(1, min(ConfigB200.concurrent_games, games)), sims_per_move=ConfigSelfPlay.mcts_iterations,
                       net=net, games_target=games, game_id_base=game_id_base, seed=ConfigB200.seed,
                       move_mode="philox", auto_restart=True, unroll=ConfigB200.graph_unroll,
                       max_free_sims=ConfigB200.max_free_sims, fin_capacity=max(games, 1),
                       index_move_greedy=ConfigMCTS.index_move_greedy,
                       dirichlet_noise=bool(ConfigMCTS.enable_dirichlet_noise), dirichlet_alpha=ConfigMCTS.dirichlet_noise_value,
                       dirichlet_ratio=ConfigMCTS.dirichlet_noise_ratio, eval_cache_log2=ConfigB200.eval_cache_log2)
    _live["runner"] = r
    return r


def play(run_id: str, plays_inferences: Optional[Dict[str, Tuple[np.ndarray, float]]] = None,
         iteration: int = 0) -> Tuple[np.ndarray, np.ndarray, np.ndarray, list]:
    """One self-play iteration on the GPU(s).  Under torchrun the ConfigB200.games_per_iteration games are sharded
    across the ranks by id (az_b200.dist.shard_games); every rank plays its shard, the finished-game records are gathered
    to rank 0 over NCCL (az_b200.dist.gather_records) and rank 0 returns all samples - the other ranks return empty
    arrays.  plays_inferences is accepted for signature compatibility: the memo lives on the device
    (ConfigB200.eval_cache_log2) and, like the reference's, never changes results."""
    from az_b200 import dist as azdist
    from az_b200.selfplay import decode_samples

    games = ConfigB200.games_per_iteration
    if ConfigGeneral.game == "chess":
        return _play_chess(run_id, games, iteration)
    rank, world = _dist()
    first, mine = azdist.shard_games(games, rank, world)
    runner = _runner(best_saved_model(run_id), mine, game_id_base=iteration * games + first)
    if mine:
        runner.run_until_done()
    fin = {k: v.contiguous() for k, v in runner.finished_device().items()}
    if world > 1:
        fin = azdist.gather_records(fin, dst=0)
    runner.fin_clear()
    rules = runner.rules
    if fin is None:  # not the collecting rank
        return (np.zeros((0, rules.height, rules.width, 4), np.float32), np.zeros((0, rules.n_actions)), np.zeros(0, np.int64), [])
    states, policies, rewards, to_end = decode_samples(rules, fin, with_distance=True)
    if ConfigSelfPlay.discounting_factor != 1:  # self_play.py:77-78: value * gamma ** (plies to the end of the game)
        rewards = rewards * ConfigSelfPlay.discounting_factor ** to_end
    return states, policies, rewards, []


def _play_chess(run_id: str, games: int, iteration: int):
    """The same iteration for chess: ChessSelfPlayRunner (az_chess_step / az_chess_move around the bf16 net); states
    float32 [S, 8, 8, 118], policies float64 [S, 1880], rewards int [S] of the finished games."""
    from az_b200.chess_selfplay import ChessSelfPlayRunner

    _live.clear()
    r = ChessSelfPlayRunner(n_trees=min(ConfigB200.concurrent_games, games), sims_per_move=ConfigSelfPlay.mcts_iterations,
                            net=best_saved_model(run_id), games_target=games, game_id_base=iteration * games,
                            seed=ConfigB200.seed, move_mode="philox", auto_restart=True, unroll=ConfigB200.graph_unroll,
                            max_free_sims=ConfigB200.max_free_sims or 8, index_move_greedy=ConfigMCTS.index_move_greedy,
                            max_plies=ConfigB200.chess_max_plies)
    _live["runner"] = r
    states, policies, rewards, known = r.run_until_done()
    return states[known], policies[known], rewards[known].astype(np.int64), []


def main(max_iterations: Optional[int] = None):
    """`python -m custom_alphazero.self_play` (one GPU) or `torchrun --nproc-per-node N -m custom_alphazero.self_play`
    (one process per GPU): the reference's loop (self_play.py:122-188).  Rank 0 talks to the serving process, writes
    samples.npz and appends to the trainer's queue; the run id and the best-model hash are broadcast so that every rank
    loads the same checkpoint, and the device memo of evaluations dies with each iteration's runner."""
    _init_distributed()
    rank, world = _dist()
    plays_inferences = reset_plays_inferences_dict()
    run_id = get_run_id() if rank == 0 else None
    if rank == 0 and run_id is None:
        run_id = time.strftime("standalone_%Y%m%d_%H%M%S")
        print(f"No serving process reachable: running stand-alone with id={run_id}")
    if world > 1:
        import torch.distributed as dist

        box = [run_id]
        dist.broadcast_object_list(box, src=0)
        run_id = box[0]
    if rank == 0:
        print(f"Starting self play with id={run_id}" + (f" on {world} ranks" if world > 1 else ""))
    iteration, previous_hash = 0, None
    while max_iterations is None or iteration < max_iterations:
        t0 = time.time()
        if rank == 0:
            os.makedirs(paths.get_self_play_iteration_path(run_id, iteration), exist_ok=True)
        current_hash = best_saved_model_hash(run_id)
        if world > 1:  # every rank must play with the checkpoint rank 0 sees
            import torch.distributed as dist

            box = [current_hash]
            dist.broadcast_object_list(box, src=0)
            current_hash = box[0]
        if previous_hash != current_hash:
            plays_inferences, previous_hash = reset_plays_inferences_dict(), current_hash
        states, policies, rewards, _ = play(run_id, plays_inferences, iteration)
        if rank == 0:
            if ConfigSelfPlay.exclude_null_games:
                keep = rewards != 0
                states, policies, rewards = states[keep], policies[keep], rewards[keep]
            print(f"Collected {len(states)} samples in {time.time() - t0:.2f} seconds")
            if (iteration + 1) % ConfigSelfPlay.samples_checkpoint_frequency == 0:
                np.savez(paths.get_self_play_samples_path(run_id, iteration), states=states, policies=policies, values=rewards)
            if not append_queue(states, policies, rewards):
                print("append_queue failed (no serving process, or it timed out): the samples of this iteration stay in "
                      "samples.npz only")
        iteration += 1
    if world > 1:
        import torch.distributed as dist

        dist.barrier()


if __name__ == "__main__":
    main(int(os.environ["AZ_MAX_ITERATIONS"]) if "AZ_MAX_ITERATIONS" in os.environ else None)
