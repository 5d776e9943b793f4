#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.json

How the reference is driven:
  * every case runs in its own subprocess with PYTHONPATH=/root/reference and a cwd
    outside this repo, because the drop-in package in this repo has the same top-level
    name (custom_alphazero) as the reference;
  * custom_alphazero.model.tensorflow.model is stubbed (TensorFlow is absent; the
    reference imports it only for a type annotation, mcts/mcts.py:9);
  * the evaluator is injected by assigning custom_alphazero.mcts.mcts.infer_sample and
    passing model=None (mcts/mcts.py:138-141) - float64 priors, Python-float value;
  * ConfigConnectN is monkey-patched per case before the first Board() (board.py:14-26).

Nothing in tests/, bench.py or the product imports this file; the JSON it writes is the
fixture.
"""
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"

MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------------------
# evaluator definitions shared (by specification, not by import) with oracle/ and the
# CUDA kernels.  State is the reference's full_state [H, W, 4] float32.
# --------------------------------------------------------------------------------------
def hash_evaluator(state, n_actions):
    """Deterministic, state-dependent evaluator with float64-exact outputs.

    cells (row-major, row 0 = top): 0 empty, 1 side-to-move stone, 2 opponent stone.
    h = FNV-1a-64 over (cell + 1); prior_a = ((mix(h, a) >> 40) % 1000 + 1) as float64
    (unnormalised), value = ((h >> 20) % 2001 - 1000) / 1000.
    """
    import numpy as np

    cells = np.argmax(state[:, :, :3], axis=-1).ravel()
    h = 0xCBF29CE484222325
    for c in cells:
        h = ((h ^ (int(c) + 1)) * 0x100000001B3) & MASK64
    priors = np.empty(n_actions, dtype=np.float64)
    for a in range(n_actions):
        m = ((h ^ ((a * 0x9E3779B97F4A7C15) & MASK64)) * 0xFF51AFD7ED558CCD) & MASK64
        priors[a] = float(((m >> 40) % 1000) + 1)
    value = (float((h >> 20) % 2001) - 1000.0) / 1000.0
    return priors, value


def worker(case):
    """Runs inside the subprocess: imports the reference and produces one case."""
    import types

    import numpy as np

    stub = types.ModuleType("custom_alphazero.model.tensorflow.model")

    class PolicyValueModel:  # annotation only
        pass

    stub.PolicyValueModel = PolicyValueModel
    sys.modules["custom_alphazero.model.tensorflow.model"] = stub

    from custom_alphazero.config import ConfigConnectN, ConfigMCTS

    ConfigConnectN.board_width = case["W"]
    ConfigConnectN.board_height = case["H"]
    ConfigConnectN.n = case["n"]
    ConfigConnectN.gravity = case["gravity"]
    assert ConfigMCTS.exploration_constant == 1.5
    assert ConfigMCTS.enable_dirichlet_noise is False

    from custom_alphazero.connect_n.board import Board
    from custom_alphazero.connect_n.move import Move

    kind = case["kind"]
    out = dict(case)

    if kind == "env_playouts":
        # SURVEY 8c "environment fingerprints": LCG-driven random playouts on Board only
        sha = hashlib.sha256()
        wins = draws = plies = 0
        detail = []
        for g in range(case["games"]):
            s = (g * 0x9E3779B97F4A7C15 + 1) & MASK64
            board = Board()
            picked = []
            while not board.is_game_over():
                moves = board.moves
                s = (s * 6364136223846793005 + 1442695040888963407) & MASK64
                idx = (s >> 33) % len(moves)
                picked.append(int(idx))
                board.play(moves[idx], keep_same_player=True)
            res = board.get_result(keep_same_player=True)
            wins += res == 1
            draws += res == 0
            plies += len(picked)
            line = "{}|{}|{}\n".format(",".join(map(str, picked)), res, repr(board))
            sha.update(line.encode())
            if g < case.get("detail_games", 0):
                detail.append(
                    {
                        "picked": picked,
                        "result": int(res),
                        "repr": repr(board),
                        "full_state_channel_sums": [
                            float(x) for x in board.full_state.sum(axis=(0, 1))
                        ],
                    }
                )
        out.update(
            wins=int(wins), draws=int(draws), plies=int(plies),
            sha16=sha.hexdigest()[:16], detail=detail,
        )
        return out

    import custom_alphazero.mcts.mcts as ref_mcts

    all_moves = Board.get_all_possible_moves()
    A = len(all_moves)
    calls = [0]

    if case["evaluator"] == "uniform":
        def evaluator(state, concurrency):
            calls[0] += 1
            return np.full(A, 1 / A), 0.0
    elif case["evaluator"] == "hash":
        def evaluator(state, concurrency):
            calls[0] += 1
            return hash_evaluator(state, A)
    else:
        raise ValueError(case["evaluator"])
    ref_mcts.infer_sample = evaluator

    def edge_stats(node):
        return {
            "actions": [all_moves.index(e.action) for e in node.edges],
            "N": [int(e.visit_count) for e in node.edges],
            "W": [float(e.total_action_value) for e in node.edges],
            "P": [float(e.prior) for e in node.edges],
        }

    if kind == "search_once":
        board = Board()
        for a in case.get("prefix", []):
            board.play(all_moves[a], keep_same_player=True)
        mcts = ref_mcts.MCTS(board, all_moves, False, {}, model=None)
        mcts.search(case["sims"])
        out.update({"edge_" + k: v for k, v in edge_stats(mcts.current_root).items()})
        out["evaluator_calls"] = calls[0]
        out["repr"] = repr(board)
        out["fullmove_number"] = int(board.fullmove_number)
        out["turn"] = int(board.turn)
        fs = board.full_state
        out["full_state_channel_sums"] = [float(x) for x in fs.sum(axis=(0, 1))]
        out["children_terminal"] = [
            bool(e.child.board.is_game_over()) for e in mcts.current_root.edges
        ]
        return out

    if kind == "full_game":
        if case.get("seed") is not None:
            np.random.seed(case["seed"])
            # one random_sample() per non-deterministic play(): record the stream the
            # engine must be fed to reproduce the game
            out["uniforms"] = [
                float(u) for u in np.random.RandomState(case["seed"]).random_sample(128)
            ]
        mcts = ref_mcts.MCTS(Board(), all_moves, False, {}, model=None)
        plies = []
        trace = ""
        states_sha = hashlib.sha256()
        while not mcts.board.is_game_over():
            mcts.search(case["sims"])
            greedy = mcts.board.fullmove_number >= ConfigMCTS.index_move_greedy
            root = mcts.current_root
            stats = edge_stats(root)
            parent_state, child_state, policy, move = mcts.play(
                greedy, return_details=True, deterministic=case.get("seed") is None
            )
            stats["move"] = all_moves.index(move)
            stats["move_str"] = str(move)
            stats["policy"] = [float(p) for p in policy]
            stats["greedy"] = bool(greedy)
            plies.append(stats)
            states_sha.update(parent_state.tobytes())
            trace += "{}:{};".format(move, ",".join(map(str, stats["N"])))
        out["plies"] = plies
        out["n_plies"] = len(plies)
        out["result"] = int(mcts.board.get_result(keep_same_player=True))
        out["sha16"] = hashlib.sha256(trace.encode()).hexdigest()[:16]
        out["states_sha16"] = states_sha.hexdigest()[:16]
        out["final_repr"] = repr(mcts.board)
        out["evaluator_calls"] = calls[0]
        return out

    raise ValueError(kind)


CASES = {
    # ---- environment fingerprints (last-but-one row) ----
    "env_6x7_n4_g": dict(kind="env_playouts", W=7, H=6, n=4, gravity=True, games=2000, detail_games=24),
    "env_9x9_n5_g": dict(kind="env_playouts", W=9, H=9, n=5, gravity=True, games=500, detail_games=24),
    "env_9x9_n5_ng": dict(kind="env_playouts", W=9, H=9, n=5, gravity=False, games=300, detail_games=8),
    "env_3x3_n3_ng": dict(kind="env_playouts", W=3, H=3, n=3, gravity=False, games=200, detail_games=24),
    "env_5x4_n3_g": dict(kind="env_playouts", W=5, H=4, n=3, gravity=True, games=300, detail_games=24),
    "env_8x8_n4_ng": dict(kind="env_playouts", W=8, H=8, n=4, gravity=False, games=200, detail_games=8),
    # ---- single searches ----
    "search_6x7_250_uniform": dict(kind="search_once", W=7, H=6, n=4, gravity=True, sims=250, evaluator="uniform"),
    "search_6x7_term_a": dict(kind="search_once", W=7, H=6, n=4, gravity=True, sims=50, evaluator="uniform", prefix=[0, 1, 0, 1, 0, 1]),
    "search_6x7_term_b": dict(kind="search_once", W=7, H=6, n=4, gravity=True, sims=400, evaluator="uniform", prefix=[0, 1, 0, 1, 0]),
    "search_6x7_800_hash": dict(kind="search_once", W=7, H=6, n=4, gravity=True, sims=800, evaluator="hash"),
    "search_6x7_1_uniform": dict(kind="search_once", W=7, H=6, n=4, gravity=True, sims=1, evaluator="uniform"),
    "search_6x7_2_uniform": dict(kind="search_once", W=7, H=6, n=4, gravity=True, sims=2, evaluator="uniform"),
    # ---- full deterministic games (play(greedy=ply>=8, deterministic=True)) ----
    "game_6x7_250_uniform": dict(kind="full_game", W=7, H=6, n=4, gravity=True, sims=250, evaluator="uniform"),
    "game_6x7_800_uniform": dict(kind="full_game", W=7, H=6, n=4, gravity=True, sims=800, evaluator="uniform"),
    "game_9x9_200_uniform": dict(kind="full_game", W=9, H=9, n=5, gravity=True, sims=200, evaluator="uniform"),
    "game_9x9ng_100_uniform": dict(kind="full_game", W=9, H=9, n=5, gravity=False, sims=100, evaluator="uniform"),
    "game_6x7_250_hash": dict(kind="full_game", W=7, H=6, n=4, gravity=True, sims=250, evaluator="hash"),
    "game_6x7_800_hash": dict(kind="full_game", W=7, H=6, n=4, gravity=True, sims=800, evaluator="hash"),
    "game_9x9_200_hash": dict(kind="full_game", W=9, H=9, n=5, gravity=True, sims=200, evaluator="hash"),
    "game_5x5ng_n3_60_hash": dict(kind="full_game", W=5, H=5, n=3, gravity=False, sims=60, evaluator="hash"),
    "game_9x9ng_100_hash": dict(kind="full_game", W=9, H=9, n=5, gravity=False, sims=100, evaluator="hash"),
    "game_3x3ng_n3_40_hash": dict(kind="full_game", W=3, H=3, n=3, gravity=False, sims=40, evaluator="hash"),
    # ---- stochastic games (np.random.seed then play(greedy, return_details=True)) ----
    "game_6x7_250_uniform_seed1234": dict(kind="full_game", W=7, H=6, n=4, gravity=True, sims=250, evaluator="uniform", seed=1234),
    "game_6x7_250_hash_seed7": dict(kind="full_game", W=7, H=6, n=4, gravity=True, sims=250, evaluator="hash", seed=7),
    "game_6x7_120_hash_seed99": dict(kind="full_game", W=7, H=6, n=4, gravity=True, sims=120, evaluator="hash", seed=99),
}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--worker":
        case = json.loads(sys.argv[2])
        print("@@RESULT@@" + json.dumps(worker(case)))
        return
    assert os.path.isdir(REFERENCE), "the reference is only mounted in the build container"
    names = sys.argv[1:] or list(CASES)
    env = dict(os.environ, PYTHONPATH=REFERENCE, PYTHONDONTWRITEBYTECODE="1")
    procs = {}
    for name in names:
        case = dict(CASES[name], name=name)
        procs[name] = subprocess.Popen(
            [sys.executable, os.path.abspath(__file__), "--worker", json.dumps(case)],
            cwd="/tmp", env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
        )
    for name, p in procs.items():
        so, se = p.communicate()
        if p.returncode != 0:
            print(se, file=sys.stderr)
            raise SystemExit(f"case {name} failed")
        payload = [l for l in so.splitlines() if l.startswith("@@RESULT@@")][0][10:]
        res = json.loads(payload)
        with open(os.path.join(HERE, name + ".json"), "w") as fp:
            json.dump(res, fp, separators=(",", ":"))
        brief = {k: res[k] for k in ("sha16", "n_plies", "result", "wins", "draws", "plies") if k in res and not isinstance(res[k], list)}
        print(name, brief)


if __name__ == "__main__":
    main()
