import ctypes, os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/custom-alphazero_b200")
import torch
from az_b200 import native
lib = native.lib(); P = lambda t: ctypes.c_void_p(t.data_ptr())
for n, cells in ((4096, 64), (4096, 42)):
    x = torch.randn(n, cells, 128, device="cuda").to(torch.bfloat16); w = torch.randn(3, 128, device="cuda"); b = torch.randn(3, device="cuda")
    out = torch.empty(n, cells, 3, device="cuda")
    def call(): native.check(lib.az_net_head_convs(P(x), P(w), P(b), n, cells, 128, P(out), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for _ in range(3): call()
    big = torch.empty(64 << 20, device="cuda")
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0
    for _ in range(10):
        big.zero_()  # flush L2
        a.record(); call(); e.record(); torch.cuda.synchronize(); tot += a.elapsed_time(e)
    ref = torch.relu(x.float() @ w.t() + b)
    print(n, cells, "%.1f us cold" % (tot / 10 * 1e3), "GB/s %.0f" % (x.numel() * 2 / (tot / 10 * 1e-3) / 1e9), "err", (out - ref).abs().max().item())
