"""Exploration: fused advance kernel vs the three-kernel route, per-phase CUDA events in steady state."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200 import engine, selfplay, net as N
from az_b200.engine import _ptr, _stream
from az_b200.native import lib, check
rules = engine.Rules(7, 6, 4, True)
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
for fused in (True, False):
    r = selfplay.SelfPlayRunner(rules, n_trees=4096, sims_per_move=800, net=N.PolicyValueNet(), games_target=1 << 40, unroll=8,
                                groups=1, max_free_sims=8, fin_capacity=16384, fused=fused)
    r.run(9600); torch.cuda.synchronize()
    g = r.groups[0]
    ev = [torch.cuda.Event(True) for _ in range(5)]
    n = 100
    acc = [0.0] * 4
    for _ in range(n):
        if fused:
            src = g.tower_carry
            hw = r.net._heads_arg()
            ev[0].record()
            check(lib().az_advance_fused(g.engine._h, _ptr(src), ctypes.byref(hw), _ptr(r.net.stem_w32), _ptr(r.net.stem_b32), _ptr(g.stem_out), _ptr(g.valid), _stream()))
            ev[1].record()
            out = r.net.tower(g.stem_out)
            ev[2].record()
            g.tower_carry.copy_(out)
            ev[3].record(); ev[4].record()
        else:
            ev[0].record(); g.engine.step(g.priors, g.values, g.states, g.valid)
            ev[1].record(); r.net(g.states, g.priors, g.values)
            ev[2].record(); g.engine.play()
            ev[3].record(); ev[4].record()
        torch.cuda.synchronize()
        for i in range(4): acc[i] += ev[i].elapsed_time(ev[i + 1])
    print(f"fused={fused}: phases us", [round(a / n * 1e3, 1) for a in acc])
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    r.run(64); torch.cuda.synchronize()
    a.record(); k = r.run(800); b.record(); torch.cuda.synchronize()
    print(f"    graph: {a.elapsed_time(b)/k*1e3:.1f} us/advance")
    del r, g
    torch.cuda.empty_cache()
