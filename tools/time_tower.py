#!/usr/bin/env python3
"""Sustained timing of az_net_tower variants (AZ_TOWER_DEBUG experiments) and the cuDNN tower with the SM clock sampled
through NVML during each timed loop.  python tools/time_tower.py [n] [debug flags ...]"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "custom-alphazero_b200")):
    sys.path.insert(0, p)
import pynvml  # noqa: E402
import torch  # noqa: E402

from az_b200 import net  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
flags = [int(v) for v in sys.argv[2:]] or [0]
torch.manual_seed(1)
inf = net.InferenceNet(net.randomise_bn(net.PolicyValueNet(6, 7, 7)))
xs = [torch.rand(n, 6, 7, 128, device="cuda").to(torch.bfloat16) for _ in range(4)]
torch.backends.cudnn.benchmark = True
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


def timed(fn, seconds=1.2):
    for i in range(10):
        fn(i % 4)
    torch.cuda.synchronize()
    t0 = time.time()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(20):
        fn(i % 4)
    torch.cuda.synchronize()
    reps = max(50, int(seconds / ((time.time() - t0) / 20)))
    clocks, power, stop = [], [], False

    def sample():
        while not stop:
            clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.05)

    th = threading.Thread(target=sample)
    th.start()
    a.record()
    for i in range(reps):
        fn(i % 4)
    b.record()
    torch.cuda.synchronize()
    stop = True
    th.join()
    clocks.sort()
    return {"ms": a.elapsed_time(b) / reps, "reps": reps, "sm_mhz_median": clocks[len(clocks) // 2], "sm_mhz_min": clocks[0],
            "power_w_max": max(power)}


out = {}
for f in flags:
    os.environ["AZ_TOWER_DEBUG"] = str(f)
    out[f"fused_debug{f}"] = timed(lambda i: inf.tower(xs[i]))
    print(f, json.dumps(out[f"fused_debug{f}"]), flush=True)
os.environ["AZ_TOWER_DEBUG"] = "0"
out["cudnn"] = timed(lambda i: inf.tower_library(xs[i]))
print("cudnn", json.dumps(out["cudnn"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "time_tower.json"), "w") as fp:
    json.dump(out, fp, indent=1)

# ---- where do the cycles go?  per-CTA counters written by the kernel itself (az_net_tower_timing)
from az_b200 import native  # noqa: E402
from az_b200.engine import _ptr  # noqa: E402

buf = torch.zeros(148 * 8 + 3 * 2048, dtype=torch.int64, device="cuda")
native.check(native.lib().az_net_tower_timing(_ptr(buf)))
for f in flags:
    os.environ["AZ_TOWER_DEBUG"] = str(f)
    inf.tower(xs[0])
    torch.cuda.synchronize()
    stamps = buf[148 * 8:].cpu().numpy().reshape(3, 2048)
    b = buf[:148 * 8].reshape(148, 8).cpu().numpy().astype(float)
    if f & 64:
        import numpy as np
        emp, iss, rdy = stamps
        nst = int((iss > 0).sum())
        d_iss = np.diff(iss[:nst])
        print("stamps", json.dumps({"stages": nst, "issue_period_median": float(np.median(d_iss)), "issue_period_mean": float(d_iss.mean()),
              "ready_frac": float(rdy[:nst].mean()),
              "issue_s_to_empty_s+8_median": float(np.median(emp[8:nst] - iss[:nst - 8])),
              "empty_s_to_issue_s_median": float(np.median(iss[8:nst] - emp[8:nst])),
              "issue_deltas_300": d_iss[300:340].tolist(), "ready_300": rdy[300:340].tolist(),
              "empty_deltas_first24": np.diff(emp[:25]).tolist(), "empty_deltas_300": np.diff(emp[300:341]).tolist(),
              "issue_minus_empty_300": (iss[300:340] - emp[300:340]).tolist()}), flush=True)
    lead = b[b[:, 0] > 0]
    epi = b[b[:, 3] > 0]
    rep = {"mma_warp_cycles": lead[:, 0].mean(), "mma_wait_act": lead[:, 1].mean(), "mma_wait_weights": lead[:, 2].mean(),
           "epi_wait_acc": epi[:, 3].mean(), "epi_body": epi[:, 4].mean(), "epi_tmem_load_first_half": epi[:, 5].mean(),
           "epi_first_half_until_published": epi[:, 6].mean(), "mma_wake_after_first_publish_of_warp2": lead[:, 7].mean(),
           "ctas_with_mma": int(len(lead)), "layers_per_cta": 9.23 * 8}
    print("timing", json.dumps(rep), flush=True)
    print("per_cta_issue_cycles(total-wait_act-wait_weights)", f, [int(x[0] - x[1] - x[2]) for x in lead], flush=True)
native.check(native.lib().az_net_tower_timing(None))
