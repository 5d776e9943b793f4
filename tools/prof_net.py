#!/usr/bin/env python3
"""A few launches of az_net_forward (the whole net in one kernel) at the headline batch, for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "custom-alphazero_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from az_b200 import net  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(1)
inf = net.InferenceNet(net.randomise_bn(net.PolicyValueNet(6, 7, 7)))
x = torch.randint(0, 2, (n, 6, 7, 4), device="cuda").to(torch.bfloat16)
for _ in range(4):
    p, v = inf(x)
torch.cuda.synchronize()
print("ok", float(p.sum()), float(v.abs().mean()))
