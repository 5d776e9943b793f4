"""Generates tests/golden/chess_*.json from the chess oracle (oracle/c/chess_oracle.c through oracle/chess_ref.py).

Unlike the Connect-N fixtures (make_golden.py runs the unmodified reference), these cannot come from the reference:
its chess layer needs python-chess, which is not installed here.  They freeze the behaviour of the oracle - whose
rules are pinned on the published perft counts - so that later changes to the oracle or the kernels show up as a diff:
  chess_playouts.json   fingerprint of 300 random playouts on the self-play path (move, mirror, ...): per game the picked
                        move indices, plies, final status; sha256 over all lines
  chess_mcts_*.json     whole self-play games by the oracle's MCTS: per ply legal-move count, chosen action, visit counts
Run from the repository root:  python tests/golden/make_chess_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import chess_ref as cr  # noqa: E402


def lcg(seed):
    s = (seed * 0x9E3779B97F4A7C15 + 1) % 2 ** 64
    while True:
        s = (s * 6364136223846793005 + 1442695040888963407) % 2 ** 64
        yield s >> 33


def playouts(n_games=300, max_plies=400):
    acts = cr.all_possible_moves()
    index = {m: i for i, m in enumerate(acts)}
    lines, total_plies, ends = [], 0, {0: 0, 1: 0, 2: 0}
    for g in range(n_games):
        rng = lcg(g)
        s = cr.start_state()
        picked = []
        for _ in range(max_plies):
            st = cr.status(s)
            if st:
                break
            moves = sorted(index[m] for m in cr.legal(s))
            a = moves[next(rng) % len(moves)]
            picked.append(a)
            s = cr.push(s, acts[a], keep_same_player=True)
        st = cr.status(s)
        ends[st] += 1
        total_plies += len(picked)
        lines.append("{}|{}|{}\n".format(",".join(map(str, picked)), st, cr.to_pos(s).tolist()))
    sha = hashlib.sha256("".join(lines).encode()).hexdigest()[:16]
    return {"games": n_games, "max_plies": max_plies, "plies": total_plies, "checkmates": ends[1], "draws": ends[2],
            "unfinished": ends[0], "sha": sha, "first_game": lines[0].split("|")[0]}


def mcts(evaluator, prior_mode, sims, max_plies, greedy_idx=8):
    g = cr.mcts_game(sims=sims, evaluator=evaluator, prior_mode=prior_mode, max_plies=max_plies, greedy_idx=greedy_idx)
    return {"evaluator": evaluator, "prior_mode": prior_mode, "sims": sims, "max_plies": max_plies, "greedy_idx": greedy_idx,
            "plies": g["plies"], "result": g["result"], "k": g["k"].tolist(), "choice": g["choice"].tolist(),
            "n": [g["n"][p][: g["k"][p]].tolist() for p in range(g["plies"])],
            "act": [g["act"][p][: g["k"][p]].tolist() for p in range(g["plies"])]}


def main():
    out = {"chess_playouts": playouts(),
           "chess_mcts_hash_f64_100": mcts("hash", "f64", 100, 60),
           "chess_mcts_uniform_f32_64": mcts("uniform", "f32", 64, 24),
           "chess_mcts_hash_f32_200": mcts("hash", "f32", 200, 40)}
    for name, data in out.items():
        with open(os.path.join(HERE, name + ".json"), "w") as fp:
            json.dump(data, fp, separators=(",", ":"))
        print(name, {k: v for k, v in data.items() if not isinstance(v, list)})


if __name__ == "__main__":
    main()
