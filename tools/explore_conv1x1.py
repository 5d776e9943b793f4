"""Times the hand-written tcgen05 shortcut GEMM (az_net_conv1x1) against cuDNN's 1x1 convolution, inside the tower
context (the input was just written by another kernel, as in a real advance) and back to back."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
import torch.nn.functional as F
from az_b200 import native
torch.backends.cudnn.benchmark = True
lib = native.lib()
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # at call time: graph capture runs on its own stream
P = lambda t: ctypes.c_void_p(t.data_ptr())
for B, H, W in ((4096, 6, 7), (4096, 8, 8), (32768, 6, 7)):
    x = torch.randn(B, H, W, 128, device="cuda").to(torch.bfloat16)
    w = (torch.randn(128, 128, 1, 1, device="cuda") * 0.1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xc = x.permute(0, 3, 1, 2)
    y = torch.empty_like(x)
    rows = B * H * W
    def ours(): native.check(lib.az_net_conv1x1(P(x), P(w), rows, 128, P(y), st()))
    def cudnn(): return F.conv2d(xc, w)
    for name, fn in (("tcgen05", ours), ("cudnn", cudnn)):
        for _ in range(5): fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                x.add_(0)  # producer kernel: leaves x in L2 like the previous convolution does
                fn()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            for _ in range(20):
                x.add_(0)
        for gg in (g, g2): gg.replay()
        a.record(); g.replay(); b.record(); torch.cuda.synchronize(); t1 = a.elapsed_time(b)
        a.record(); g2.replay(); b.record(); torch.cuda.synchronize(); t2 = a.elapsed_time(b)
        print(f"{B}x{H}x{W} {name}: {(t1 - t2) / 20 * 1e3:.1f} us per call ({rows * 512 / ((t1 - t2) / 20 * 1e-3) / 1e9:.0f} GB/s algorithmic)")

# the whole tower with and without the hand-written shortcut, same process (same cuDNN algorithm choices)
from az_b200.net import InferenceNet, PolicyValueNet
for B, H, W in ((4096, 6, 7), (4096, 8, 8)):
    inf = InferenceNet(PolicyValueNet(H, W, 7), dtype=torch.bfloat16, device="cuda")
    h0 = torch.randn(B, H, W, 128, device="cuda").to(torch.bfloat16)
    res = {}
    for flag in (True, False, True, False):
        inf.tc_shortcut = flag
        for _ in range(3): inf.tower(h0)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10): inf.tower(h0)
        g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); g.replay(); b.record(); torch.cuda.synchronize()
        print(f"tower {B}x{H}x{W} tc_shortcut={flag}: {a.elapsed_time(b) / 20 * 1e3:.1f} us")
