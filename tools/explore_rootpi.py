"""Exploration: root policy targets, bf16 GPU net vs fp32 CPU net, through the compat MCTS."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import numpy as np, torch
from az_b200 import net as N
from custom_alphazero.connect_n.board import Board
from custom_alphazero.connect_n.move import Move
import custom_alphazero.mcts.mcts as m

torch.manual_seed(0)
for scale, label in [(1.0, "random init"), (6.0, "sharpened heads (x6 policy/value dense weights)")]:
    ref = N.randomise_bn(N.PolicyValueNet()).eval()
    with torch.no_grad():
        ref.policy_fc.weight.mul_(scale); ref.value_fc2.weight.mul_(scale)
    inf = N.InferenceNet(ref, device="cuda")
    class CpuModel:
        def __call__(self, x):
            with torch.no_grad():
                p, v = ref(torch.from_numpy(x))
            return p, v
    class GpuModel:
        def __call__(self, x):
            p, v = inf(torch.from_numpy(x).cuda().to(torch.bfloat16))
            return p.cpu(), v.cpu()
    all_moves = Board.get_all_possible_moves()
    worst = 0.0; diffs = []
    rng = np.random.RandomState(1)
    for pos in range(12):
        b = Board()
        for _ in range(rng.randint(0, 10)):
            mv = b.moves
            if not mv or b.is_game_over(): break
            b.play(mv[rng.randint(len(mv))], keep_same_player=True)
        if b.is_game_over(): continue
        pis = []
        for model in (CpuModel(), GpuModel()):
            t = m.MCTS(b, all_moves, False, {}, model=model)
            t.search(400)
            n = np.asarray([e.visit_count for e in t.current_root.edges], dtype=np.float64)
            pis.append(n / n.sum())
        d = np.abs(pis[0] - pis[1]).max()
        diffs.append(d)
    print(label, "max |d pi| per position:", np.round(diffs, 4).tolist(), "max", max(diffs), "mean", np.mean(diffs))
