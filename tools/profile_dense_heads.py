"""One az_net_dense_heads call at 4096 positions (for ncu) and its event timing against the library route."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200 import native
lib = native.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())
n, A = 4096, 1880
wp = torch.cat([(torch.randn(A, 128, device="cuda") * 0.5).to(torch.bfloat16), torch.zeros(40, 128, device="cuda", dtype=torch.bfloat16)]).contiguous()
bp = torch.rand(A, device="cuda"); w1 = (torch.randn(256, 64, device="cuda") * 0.2).to(torch.bfloat16).contiguous()
b1 = torch.rand(256, device="cuda"); w2 = torch.randn(256, device="cuda"); b2 = torch.tensor([0.1], device="cuda")
hd = torch.relu(torch.randn(n, 64, 3, device="cuda")).contiguous()
priors = torch.empty(n, A, device="cuda"); values = torch.empty(n, device="cuda"); scratch = torch.empty(n, native.AZ_DENSE_HEAD_SPLITS, 2, device="cuda")
def call():
    native.check(lib.az_net_dense_heads(P(hd), P(wp), P(bp), P(w1), P(b1), P(w2), P(b2), n, 64, A, P(priors), P(values), P(scratch),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
for _ in range(3): call()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): call()
b.record(); torch.cuda.synchronize()
print(f"az_net_dense_heads: {a.elapsed_time(b) / 20 * 1e3:.1f} us per call")
