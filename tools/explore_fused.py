"""Exploration: cuDNN fused conv+bias+relu / conv+add+relu ops vs separate elementwise kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch, torch.nn.functional as F

def ev_time(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

torch.backends.cudnn.benchmark = True
T = 4096
for dtype in (torch.bfloat16, torch.float16):
    x = torch.randn(T, 128, 6, 7, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
    w3 = (torch.randn(128, 128, 3, 3, device="cuda", dtype=dtype) * 0.03).contiguous(memory_format=torch.channels_last)
    w1 = (torch.randn(128, 128, 1, 1, device="cuda", dtype=dtype) * 0.1).contiguous(memory_format=torch.channels_last)
    b = torch.randn(128, device="cuda", dtype=dtype)
    z = torch.randn_like(x)
    print(dtype)
    print("  conv3x3 nobias         %.1f us" % ev_time(lambda: F.conv2d(x, w3, None, padding=1)))
    print("  conv3x3 bias           %.1f us" % ev_time(lambda: F.conv2d(x, w3, b, padding=1)))
    print("  conv3x3 bias relu_     %.1f us" % ev_time(lambda: F.relu_(F.conv2d(x, w3, b, padding=1))))
    print("  conv1x1 nobias         %.1f us" % ev_time(lambda: F.conv2d(x, w1, None)))
    try:
        f = lambda: torch.cudnn_convolution_relu(x, w3, b, (1, 1), (1, 1), (1, 1), 1)
        y = f(); ref = F.relu(F.conv2d(x, w3, b, padding=1))
        print("  cudnn_convolution_relu %.1f us  maxdiff %.4f cl=%s" % (ev_time(f), (y.float() - ref.float()).abs().max().item(), y.is_contiguous(memory_format=torch.channels_last)))
    except Exception as e:
        print("  cudnn_convolution_relu failed:", repr(e)[:200])
    try:
        f = lambda: torch.cudnn_convolution_add_relu(x, w3, z, 1.0, b, (1, 1), (1, 1), (1, 1), 1)
        y = f(); ref = F.relu(F.conv2d(x, w3, b, padding=1) + z)
        print("  cudnn_conv_add_relu    %.1f us  maxdiff %.4f" % (ev_time(f), (y.float() - ref.float()).abs().max().item()))
    except Exception as e:
        print("  cudnn_convolution_add_relu failed:", repr(e)[:200])
    xm = x.permute(0, 2, 3, 1).reshape(-1, 128)
    wm = w1.reshape(128, 128)
    print("  linear 1x1 (cublasLt bias epilogue) %.1f us" % ev_time(lambda: F.linear(xm, wm, b)))
    y = F.conv2d(x, w3, None, padding=1)
    print("  y + b[None,:,None,None] %.1f us" % ev_time(lambda: y + b[None, :, None, None]))
    ym = y.permute(0, 2, 3, 1).reshape(-1, 128)
    print("  ym + b  (2-D)           %.1f us" % ev_time(lambda: ym + b))
    print("  relu_                   %.1f us" % ev_time(lambda: F.relu_(y)))
    print("  add_                    %.1f us" % ev_time(lambda: y.add_(z)))
