#!/usr/bin/env python3
"""az_net_tower (one persistent tcgen05 kernel) against the cuDNN tower it replaces: a correctness summary that says
WHERE a mismatch is (position in tile, board row / column, channel) and CUDA-event timings at the headline batch.
Writes gpurun_out/tower_timing.json.   python tools/explore_tower.py [n_positions]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "custom-alphazero_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from az_b200 import engine, native, net  # noqa: E402


def reference(x, blocks):
    t = x.float().permute(0, 3, 1, 2)
    for w1, b1, w2, wp, b2p in blocks:
        h = F.relu(F.conv2d(t, w1, b1, padding=1)).to(torch.bfloat16).float()
        t = F.relu(F.conv2d(h, w2, b2p, padding=1) + F.conv2d(t, wp)).to(torch.bfloat16).float()
    return t.permute(0, 2, 3, 1)


def run(xd, img, bias, H, W, depth, out=None):
    out = torch.empty_like(xd) if out is None else out
    native.check(native.lib().az_net_tower(engine._ptr(xd), engine._ptr(img), engine._ptr(bias), xd.shape[0], H, W, 128,
                                           depth, 0, engine._ptr(out), engine._stream()))
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    H, W, depth = 6, 7, 4
    dev = "cuda"
    report = {}
    g = torch.Generator().manual_seed(0)
    blocks = []
    for _ in range(depth):
        w1 = (torch.randn(128, 128, 3, 3, generator=g) * 0.03).to(torch.bfloat16).float()
        w2 = (torch.randn(128, 128, 3, 3, generator=g) * 0.03).to(torch.bfloat16).float()
        wp = (torch.randn(128, 128, 1, 1, generator=g) * 0.08).to(torch.bfloat16).float()
        blocks.append((w1, torch.randn(128, generator=g) * 0.1, w2, wp, torch.randn(128, generator=g) * 0.1))
    img, bias = net.pack_tower_weights(blocks)
    img, bias = img.to(dev), bias.to(dev)

    # ---- correctness, small: where is the error?
    for d in (1, 4):
        x = torch.rand(10, H, W, 128, generator=g).to(torch.bfloat16)
        got = run(x.to(dev), img[: d * 38 * 8192], bias[:d].contiguous(), H, W, d).float().cpu()
        torch.cuda.synchronize()
        want = reference(x, blocks[:d])
        err = (got - want).abs()
        info = {"max_err": float(err.max()), "max_ref": float(want.abs().max()), "nan": int(torch.isnan(got).sum()),
                "frac_gt_1e-2": float((err > 1e-2).float().mean())}
        if info["max_err"] > 0.05 or info["nan"]:
            bad = err > 0.05
            info["bad_by_position"] = bad.float().mean(dim=(1, 2, 3)).tolist()
            info["bad_by_y"] = bad.float().mean(dim=(0, 2, 3)).tolist()
            info["bad_by_x"] = bad.float().mean(dim=(0, 1, 3)).tolist()
            info["bad_by_channel_16"] = bad.float().mean(dim=(0, 1, 2)).reshape(8, 16).mean(1).tolist()
            info["sample_got"] = got[0, 0, 0, :8].tolist()
            info["sample_want"] = want[0, 0, 0, :8].tolist()
        report[f"check_depth{d}"] = info
        print(f"depth {d}:", json.dumps(info), flush=True)

    # ---- timing at n positions: inputs rotate through > L2 worth of buffers
    torch.manual_seed(1)
    fp32 = net.randomise_bn(net.PolicyValueNet(H, W, 7))
    inf = net.InferenceNet(fp32)
    n_buf = 4
    xs = [torch.rand(n, H, W, 128, device=dev).to(torch.bfloat16) for _ in range(n_buf)]
    outs = [torch.empty_like(xs[0]) for _ in range(n_buf)]
    torch.backends.cudnn.benchmark = True
    burn = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    for _ in range(100):
        burn @ burn
    torch.cuda.synchronize()

    def time_it(fn, reps=40):
        for i in range(5):
            fn(i % n_buf)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i % n_buf)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    fused_ms = time_it(lambda i: run(xs[i], inf.tower_img, inf.tower_bias, H, W, depth, outs[i]))
    lib_ms = time_it(lambda i: inf.tower_library(xs[i]))
    # under a CUDA graph (how the self-play loop runs it)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(n_buf):
            inf.tower_library(xs[i])
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    lib_graph_ms = a.elapsed_time(b) / (10 * n_buf)
    flops = n * 4 * (2 * 2 * 42 * 9 * 128 * 128 + 2 * 42 * 128 * 128)
    diff = (run(xs[0], inf.tower_img, inf.tower_bias, H, W, depth).float() - inf.tower_library(xs[0]).float()).abs()
    report["timing"] = {"positions": n, "fused_ms": fused_ms, "cudnn_ms": lib_ms, "cudnn_graph_ms": lib_graph_ms,
                        "fused_tflops": flops / fused_ms / 1e9, "cudnn_tflops": flops / lib_graph_ms / 1e9,
                        "max_abs_diff_fused_vs_cudnn": float(diff.max())}
    print(json.dumps(report["timing"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tower_timing.json"), "w") as fp:
        json.dump(report, fp, indent=1)


if __name__ == "__main__":
    main()
