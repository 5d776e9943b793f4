"""Exploration: where does an advance spend its time in steady state? (eager, per-phase CUDA events)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200 import engine, selfplay, net as N
rules = engine.Rules(7, 6, 4, True)
torch.manual_seed(0)
fp32 = N.PolicyValueNet()
torch.backends.cudnn.benchmark = True
for T, mf in [(4096, 8), (4096, 1), (2048, 8)]:
    r = selfplay.SelfPlayRunner(rules, n_trees=T, sims_per_move=800, net=fp32, games_target=1 << 40, unroll=8,
                                groups=1, max_free_sims=mf, fin_capacity=16384)
    r.run(800 * 12); torch.cuda.synchronize()
    g = r.groups[0]
    ev = [torch.cuda.Event(True) for _ in range(4)]
    acc = [0.0, 0.0, 0.0]
    n = 200
    for _ in range(n):
        ev[0].record(); g.engine.step(g.priors, g.values, g.states, g.valid)
        ev[1].record(); r.net(g.states, g.priors, g.values)
        ev[2].record(); g.engine.play()
        ev[3].record(); torch.cuda.synchronize()
        for i in range(3): acc[i] += ev[i].elapsed_time(ev[i + 1])
    print(f"T={T} max_free={mf}: k_step {acc[0]/n*1e3:.1f} us  net {acc[1]/n*1e3:.1f} us  k_play {acc[2]/n*1e3:.1f} us", flush=True)
    # graph advance for comparison
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    r.run(64); torch.cuda.synchronize()
    a.record(); k = r.run(800); b.record(); torch.cuda.synchronize()
    print(f"    graph: {a.elapsed_time(b)/k*1e3:.1f} us/advance; max nodes used {int(g.engine.view('n_nodes').max())} cap {g.engine.cfg.node_capacity}")
    del r, g
    torch.cuda.empty_cache()
