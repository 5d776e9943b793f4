"""Exploration: where does the e2e step lose time vs the device-only step?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200 import engine, selfplay, net as N
rules = engine.Rules(7, 6, 4, True)
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
r = selfplay.SelfPlayRunner(rules, n_trees=4096, sims_per_move=800, net=N.PolicyValueNet(), games_target=1 << 40, unroll=8,
                            groups=1, max_free_sims=8, fin_capacity=16384)
r.run(800 * 8); torch.cuda.synchronize(); r.fin_clear()
flat_dev = r.net.flat_weights(); flat_host = flat_dev.cpu().pin_memory(); params = list(r.net.parameters())
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = T()
    flat_dev.copy_(flat_host, non_blocking=True)
    off = 0
    with torch.no_grad():
        for p_ in params:
            p_.copy_(flat_dev[off: off + p_.numel()].view_as(p_)); off += p_.numel()
    t1 = T()
    r.run(800)
    t2 = T()
    fin = r.finished_device()
    t3 = T()
    st, po, va = selfplay.decode_samples(rules, fin)
    t4 = T()
    r.fin_clear()
    t5 = T()
    print(f"weights {1e3*(t1-t0):.2f} ms | run {1e3*(t2-t1):.2f} | fin views {1e3*(t3-t2):.2f} | decode {1e3*(t4-t3):.2f} ({len(va)} samples) | clear {1e3*(t5-t4):.2f}")
# unsynchronised variant (as in bench)
torch.cuda.synchronize(); t0 = time.perf_counter()
for it in range(4):
    flat_dev.copy_(flat_host, non_blocking=True)
    off = 0
    with torch.no_grad():
        for p_ in params:
            p_.copy_(flat_dev[off: off + p_.numel()].view_as(p_)); off += p_.numel()
    r.run(800)
    fin = r.finished_device(); st, po, va = selfplay.decode_samples(rules, fin); r.fin_clear()
torch.cuda.synchronize(); print("bench-style e2e per step ms", (time.perf_counter() - t0) / 4 * 1e3)
torch.cuda.synchronize(); t0 = time.perf_counter()
for it in range(4): r.run(800)
torch.cuda.synchronize(); print("device-only per step ms", (time.perf_counter() - t0) / 4 * 1e3)
