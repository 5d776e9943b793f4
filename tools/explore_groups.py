"""Exploration: groups x max_free_sims sweep of the runner in steady state."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200 import engine, selfplay, net as N
rules = engine.Rules(7, 6, 4, True)
torch.manual_seed(0)
fp32 = N.PolicyValueNet()
torch.backends.cudnn.benchmark = True
for groups, mf, extra in [(1, 8, 0), (2, 8, 0), (2, 4, 0), (3, 8, 0)]:
    r = selfplay.SelfPlayRunner(rules, n_trees=4096, sims_per_move=800, net=fp32, games_target=1 << 40, unroll=8,
                                groups=groups, max_free_sims=mf, fin_capacity=16384, extra_sims=extra)
    r.run(800 * 12); torch.cuda.synchronize()   # 12 moves in: trees desynchronised, terminal hits appear
    t0 = r.totals()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(); n = r.run(1600); b.record(); torch.cuda.synchronize()
    t1 = r.totals(); ms = a.elapsed_time(b)
    print(f"groups={groups} max_free={mf} extra={extra}: {ms/n*1e3:.1f} us/advance  sims/s {(t1['sims']-t0['sims'])/ms*1e3/1e6:.3f} M  evals/s {(t1['evals']-t0['evals'])/ms*1e3/1e6:.3f} M", flush=True)
    r.check_status()
    del r
    torch.cuda.empty_cache()
