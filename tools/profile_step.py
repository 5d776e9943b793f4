"""Steady-state driver for ncu: 12 moves of warm-up under graphs, then a few eager advances."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200 import engine, selfplay, net as N
rules = engine.Rules(7, 6, 4, True)
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
mf = int(os.environ.get("MF", 8))
r = selfplay.SelfPlayRunner(rules, n_trees=4096, sims_per_move=800, net=N.PolicyValueNet(), games_target=1 << 40, unroll=8,
                            groups=1, max_free_sims=mf, fin_capacity=16384)
r.run(9600); torch.cuda.synchronize()
g = r.groups[0]
g.tower_out = None
for _ in range(10):
    r._advance(g)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("steady_advances")   # ncu --nvtx --nvtx-include "steady_advances/" profiles exactly these
for _ in range(2):
    r._advance(g)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done", r.totals())
