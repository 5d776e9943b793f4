"""Exploration: device timeline of az_advance_fused inside graph replays (start/end per launch)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch, numpy as np
from az_b200 import engine, selfplay, net as N
from az_b200.engine import _ptr
from az_b200.native import lib, check
rules = engine.Rules(7, 6, 4, True)
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
for mf, extra in [(8, 0), (1, 16)]:
    r = selfplay.SelfPlayRunner(rules, n_trees=4096, sims_per_move=800, net=N.PolicyValueNet(), games_target=1 << 40, unroll=8,
                                groups=1, max_free_sims=mf, fin_capacity=16384, extra_sims=extra)
    slots = torch.zeros(16, dtype=torch.int64, device="cuda")
    check(lib().az_debug_timeline(r.engine._h, _ptr(slots), 8))
    r.run(9600); torch.cuda.synchronize()
    slots.view(8, 2)[:, 0] = torch.iinfo(torch.int64).max; slots.view(8, 2)[:, 1] = 0
    r.run(8); torch.cuda.synchronize()
    tl = slots.view(8, 2).cpu().numpy().astype(np.int64)
    dur = (tl[:, 1] - tl[:, 0]) / 1e3
    gap = (tl[1:, 0] - tl[:-1, 1]) / 1e3
    print(f"max_free={mf} extra={extra}: k_advance us {np.round(dur,1).tolist()}  between (tower etc.) us {np.round(gap,1).tolist()}  period {np.round((tl[1:,0]-tl[:-1,0])/1e3,1).tolist()}")
    del r
    torch.cuda.empty_cache()
