#!/usr/bin/env python3
"""A/B timing of the tower / whole-net kernel of two builds of the package in one run on one box:
python tools/ab_tower.py <pkg_root_a> <pkg_root_b> ...   (each a directory holding az_b200/ and libaz_b200.so).
Each build runs in its own subprocess (the library and the weight packing differ), alternating, three rounds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, time
sys.path.insert(0, sys.argv[1])
import torch
from az_b200 import net
torch.manual_seed(1)
n = 4096
inf = net.InferenceNet(net.randomise_bn(net.PolicyValueNet(6, 7, 7)))
xs = [torch.rand(n, 6, 7, 128, device="cuda").to(torch.bfloat16) for _ in range(4)]
ps = [torch.randint(0, 2, (n, 6, 7, 4), device="cuda").to(torch.bfloat16) for _ in range(4)]
def timed(fn, reps=1500):
    for i in range(50): fn(i % 4)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i % 4)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = {"tower_ms": timed(lambda i: inf.tower(xs[i])), "net_ms": timed(lambda i: inf(ps[i]))}
p, v = inf(ps[0]); torch.cuda.synchronize()
out["checksum"] = [float(p.double().sum()), float(v.double().abs().sum())]
print(json.dumps(out))
'''
res = {}
for rnd in range(3):
    for root in sys.argv[1:]:
        r = subprocess.run([sys.executable, "-c", CHILD, root], capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(root, "FAILED", r.stderr[-2000:])
            continue
        res.setdefault(root, []).append(json.loads(line[-1]))
        print(rnd, root, line[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_tower.json"), "w"), indent=1)
