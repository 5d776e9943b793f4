"""Exploration: custom block tail (shortcut GEMM + add + ReLU) vs the cuDNN route."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200 import net as N
from az_b200.engine import _ptr, _stream
from az_b200.native import lib, check
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
fp32 = N.randomise_bn(N.PolicyValueNet()).eval()

def graph_time(fn, n=50):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5): g.replay()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

T = 4096
x = torch.zeros(T, 6, 7, 4); code = torch.randint(0, 3, (T, 6, 7)); x.scatter_(3, code[..., None], 1.0); x[..., 3] = 1
xd = x.cuda().to(torch.bfloat16)
outs = {}
for tail in (False, True):
    inf = N.InferenceNet(fp32, device="cuda"); inf.custom_tail = tail
    p, v = inf(xd); outs[tail] = (p.clone(), v.clone())
    print(f"custom_tail={tail}: net forward {graph_time(lambda: inf(xd)):.1f} us", flush=True)
print("max |dp|", (outs[0][0] - outs[1][0]).abs().max().item(), "max |dv|", (outs[0][1] - outs[1][1]).abs().max().item())
with torch.no_grad(): wp_, wv_ = fp32(x)
for tail in (False, True):
    print(f"custom_tail={tail} vs fp32: dp {(outs[tail][0].cpu()-wp_).abs().max().item():.5f} dv {(outs[tail][1].cpu()-wv_.reshape(-1)).abs().max().item():.5f}")
# the kernel alone
xx = torch.randn(T * 42, 128, device="cuda").to(torch.bfloat16); c2 = torch.randn(T * 42, 128, device="cuda").to(torch.bfloat16)
w = (torch.randn(128, 128, device="cuda") * 0.1).contiguous(); b = torch.randn(128, device="cuda")
want = torch.relu(c2.float() + xx.float() @ w.to(torch.bfloat16).float().T + b)
got = c2.clone()
check(lib().az_net_block_tail(_ptr(xx), _ptr(got), _ptr(w), _ptr(b), T * 42, 128, _stream()))
print("kernel vs torch: max abs diff", (got.float() - want).abs().max().item(), "rel to max", want.abs().max().item())
buf = c2.clone()
print("k_block_tail alone: %.1f us" % graph_time(lambda: check(lib().az_net_block_tail(_ptr(xx), _ptr(buf), _ptr(w), _ptr(b), T * 42, 128, _stream()))))
