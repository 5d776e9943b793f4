import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import numpy as np, torch
from az_b200 import engine, net, selfplay, train
rules = engine.Rules(7, 6, 4, True)
torch.manual_seed(0)
fp32 = net.PolicyValueNet()
runner = selfplay.SelfPlayRunner(rules, n_trees=64, sims_per_move=16, net=fp32, games_target=96, unroll=4, seed=3)
for it in range(2):
    runner.reset()
    n = runner.run_until_done()
    fin = runner.finished_device()
    print("advances", n, "games", fin["len"].numel(), "totals", runner.totals(), "active", runner.active_trees())
    s, p, v = selfplay.decode_samples(rules, fin, exclude_null_games=True)
    print("samples", len(v), "results", np.bincount(fin["result"].cpu().numpy()))
    runner.fin_clear()
