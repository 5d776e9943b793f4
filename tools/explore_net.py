"""Exploration: net-forward variants under a CUDA graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200 import net as N
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
fp32 = N.PolicyValueNet()

def graph_time(inf, x, n=50):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): inf(x)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        inf(x)
    for _ in range(5): g.replay()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

for T in (4096, 4059, 2048, 2255, 8192):
    x = torch.randint(0, 2, (T, 6, 7, 4), device="cuda").to(torch.bfloat16)
    for overlap in (False, True):
        inf = N.InferenceNet(fp32, device="cuda")
        inf.overlap_shortcut = overlap
        us = graph_time(inf, x)
        print(f"T={T} overlap_shortcut={overlap}: {us:.1f} us -> {T*N.flops_per_eval(6,7,7)/us/1e6:.0f} TFLOP/s, {us/T*1e3:.1f} ns/position", flush=True)
