"""Device timeline of the whole-net route inside graph replays: {start, end} of every az_step_gather and az_net_forward
launch of one captured graph (globaltimer stamps written by the kernels), in steady state."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import numpy as np, torch
from az_b200 import engine, selfplay, net as N
from az_b200.engine import _ptr
from az_b200.native import lib, check
rules = engine.Rules(7, 6, 4, True)
torch.manual_seed(0)
U = 8
r = selfplay.SelfPlayRunner(rules, n_trees=4096, sims_per_move=800, net=N.PolicyValueNet(), games_target=1 << 40, unroll=U,
                            max_free_sims=int(os.environ["MF"]) if "MF" in os.environ else None,
                            net_tree_sims=int(os.environ["INSIDE"]) if "INSIDE" in os.environ else None, fin_capacity=16384)
r.run(int(os.environ.get("PREROLL", "40000"))); torch.cuda.synchronize()
for g in r.groups: g.engine.fin_clear()
ts = torch.zeros(2 * U, dtype=torch.int64, device="cuda")
tn = torch.zeros(2 * U, dtype=torch.int64, device="cuda")
check(lib().az_debug_timeline(r.engine._h, _ptr(ts), U))
check(lib().az_net_debug_timeline(_ptr(tn), U))
cyc = torch.zeros((148, 8), dtype=torch.int64, device="cuda")
if os.environ.get("CYCLES"):
    check(lib().az_net_tower_timing(_ptr(cyc)))  # in-kernel cycle counters of the last launch: true SM clock = cycles / time
r.graph = None
r.capture()
def clear():
    for t in (ts, tn):
        t.view(U, 2)[:, 0] = torch.iinfo(torch.int64).max; t.view(U, 2)[:, 1] = 0
rows = []
for rep in range(6):
    clear(); torch.cuda.synchronize()
    r.run(U); torch.cuda.synchronize()
    s = ts.view(U, 2).cpu().numpy().astype(np.int64); n = tn.view(U, 2).cpu().numpy().astype(np.int64)
    cnt = int(r.groups[0].leaf_count[0])
    for i in range(U):
        rows.append({"step_us": (s[i, 1] - s[i, 0]) / 1e3, "step_to_net_us": (n[i, 0] - s[i, 1]) / 1e3, "net_us": (n[i, 1] - n[i, 0]) / 1e3,
                     "net_to_step_us": ((s[i + 1, 0] - n[i, 1]) / 1e3) if i + 1 < U else None,
                     "period_us": ((s[i + 1, 0] - s[i, 0]) / 1e3) if i + 1 < U else None, "leaves_last": cnt})
check(lib().az_debug_timeline(r.engine._h, None, 0)); check(lib().az_net_debug_timeline(None, 0))
def med(k):
    v = [x[k] for x in rows if x[k] is not None]
    return float(np.median(v)), float(np.mean(v)), float(np.min(v)), float(np.max(v))
out = {k: dict(zip(("median", "mean", "min", "max"), med(k))) for k in ("step_us", "step_to_net_us", "net_us", "net_to_step_us", "period_us")}
out["leaves_last"] = rows[-1]["leaves_last"]
if os.environ.get("CYCLES"):
    c = cyc.cpu().numpy().astype(float)
    lead = c[c[:, 0] > 0]
    out["mma_warp_cycles_last_launch_max"] = float(lead[:, 0].max())
    out["mma_warp_cycles_last_launch_mean"] = float(lead[:, 0].mean())
    out["net_us_last_launch"] = rows[-1]["net_us"]
    out["implied_sm_mhz"] = float(lead[:, 0].max()) / rows[-1]["net_us"]
    out["wait_act_mean"] = float(lead[:, 1].mean()); out["wait_weights_mean"] = float(lead[:, 2].mean())
    out["per_cta_total_wait_act_wait_weights"] = [[int(c[i, 0]), int(c[i, 1]), int(c[i, 2])] for i in range(148) if c[i, 0] > 0]
    check(lib().az_net_tower_timing(None))
out["max_free_sims"], out["tree_sims_inside_net"] = r.max_free_sims, r.net_tree_sims
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"summary": out, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", os.environ.get("OUT", "timeline2.json")), "w"), indent=1)
