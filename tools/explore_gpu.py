"""Exploratory timings on the GPU box (not part of the bench contract)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200 import engine, selfplay, net as N

def ev_time(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

rules = engine.Rules(7, 6, 4, True)
T = int(os.environ.get("T", 4096))
torch.manual_seed(0)
fp32 = N.PolicyValueNet()
for dtype in (torch.bfloat16, torch.float16):
    inf = N.InferenceNet(fp32, dtype=dtype, device="cuda")
    x = torch.randint(0, 2, (T, 6, 7, 4), device="cuda").to(dtype)
    torch.backends.cudnn.benchmark = True
    ms = ev_time(lambda: inf(x))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        inf(x)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        out = inf(x)
    msg = ev_time(lambda: g.replay())
    fl = N.flops_per_eval(6, 7, 7) * T
    print(f"net {dtype} T={T}: eager {ms:.3f} ms, graph {msg:.3f} ms -> {fl/msg/1e9:.1f} TFLOP/s, {T/msg*1e3/1e6:.2f} M evals/s")

# full runner
r = selfplay.SelfPlayRunner(rules, n_trees=T, sims_per_move=800, net=fp32, games_target=10**9, unroll=8)
r.run(64); torch.cuda.synchronize()
t0 = r.engine.totals()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record(); n = r.run(800); b.record(); torch.cuda.synchronize()
t1 = r.engine.totals()
ms = a.elapsed_time(b)
print(f"runner: {n} advances in {ms:.1f} ms = {ms/n*1e3:.1f} us/advance; sims/s {(t1['sims']-t0['sims'])/ms*1e3:.0f} evals/s {(t1['evals']-t0['evals'])/ms*1e3:.0f} moves {t1['moves']-t0['moves']}")
r.engine.check_status()
# tree kernels alone
st = r.states; va = r.valid
ms_step = ev_time(lambda: r.engine.step(r.priors, r.values, st, va), n=50)
ms_play = ev_time(lambda: r.engine.play(), n=50)
print(f"k_step {ms_step*1e3:.1f} us, k_play {ms_play*1e3:.1f} us")
# fixed evaluator search throughput
for ev in ("uniform", "hash"):
    e = engine.TreeEngine(rules, n_trees=T, sims_per_move=800, eval_mode=ev, prior_mode="f64")
    def one():
        e.search(); e.play()
    for _ in range(2): one()
    torch.cuda.synchronize()
    a.record()
    for _ in range(4): one()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 4
    print(f"fixed {ev}: search(800)+play for {T} trees: {ms:.2f} ms -> {T*800/ms*1e3/1e6:.1f} M sims/s")

# net accuracy bf16 fast vs fp32
torch.manual_seed(0)
ref = N.randomise_bn(N.PolicyValueNet()).eval()
xs = torch.zeros(2048, 6, 7, 4); code = torch.randint(0, 3, (2048, 6, 7)); xs.scatter_(3, code[..., None], 1.0); xs[..., 3] = 1
with torch.no_grad(): wp, wv = ref(xs)
inf = N.InferenceNet(ref, dtype=torch.bfloat16, device="cuda")
p, v = inf(xs.cuda().to(torch.bfloat16))
print("bf16 fast vs fp32: max dp %.5f mean dp %.6f max dv %.5f" % ((p.cpu()-wp).abs().max(), (p.cpu()-wp).abs().mean(), (v.cpu()-wv.reshape(-1)).abs().max()))
