"""Tower time against the number of positions: does cuDNN's 256x128-tile kernel pay for a nearly empty last wave?
(4096 positions x 42 cells = 672 tiles of 256 rows on 74 CTA pairs = 9.08 waves.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200.net import InferenceNet, PolicyValueNet
torch.backends.cudnn.benchmark = True
inf = InferenceNet(PolicyValueNet(6, 7, 7), dtype=torch.bfloat16, device="cuda")
for B in (4096, 4080, 4064, 4059, 4050, 4032, 3996, 3608, 4510, 8192, 8118):
    h0 = torch.randn(B, 6, 7, 128, device="cuda").to(torch.bfloat16)
    for _ in range(3): inf.tower(h0)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): inf.tower(h0)
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); g.replay(); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    print(f"B={B}: tower {us:.1f} us, {us / B * 1e3:.2f} ns per position, tiles256={B * 42 / 256:.1f} waves={B * 42 / 256 / 74:.2f}")
