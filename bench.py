#!/usr/bin/env python3
"""bench.py - Connect-4 self-play MCTS simulations/s on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's engine
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Workload (config C2 of BASELINE.json, per GPU): 4096 concurrent 6x7 Connect-4 games, 800 simulations
per move, every leaf evaluated by the bf16 policy/value net (random-init weights of the reference
architecture: synthetic), finished games refilled so the batch stays full.  One "step" = 800 lock-step
advances of all trees = one move's worth of simulations for every game (~3.3 M simulations per GPU).
N > 1: weak scaling - every rank owns 4096 games and a net replica (32768 games at N = 8 = config C3);
the only collectives are the per-step weight broadcast and, in the e2e leg, the game-record gather.

Prints ONE JSON line (rank 0).  `value` = whole-job simulations/s, device-timed with everything
resident in HBM; `e2e` = the same metric through the public API with host buffers: every step uploads
the weights from pinned host memory and downloads the decoded training samples of the games that
finished.  `roofline` is the net forward (tensor bound, the dominant kernels); `roofline_tree` the
az_step kernel (HBM bound).  `cpu_baseline` = the oracle's Python port of the reference run the way the
reference runs self-play (os.cpu_count()-1 processes, batch-1 fp32 CPU net), on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "custom-alphazero_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "connect4_selfplay_mcts_simulations_per_sec"
UNIT = "sims/s"
RULES = (7, 6, 4, True)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            d = json.load(fp)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.tmp = None

    def start(self):
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.tmp.close()
        os.unlink(self.tmp.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def run_reference(args):
    """--impl reference: the reference's CPU self-play path (oracle port, see oracle/cpu_baseline.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline

    sp = cpu_baseline.CpuSelfPlay(RULES, args.sims, "net", workers=None)
    try:
        for _ in range(max(args.warmup, 1)):
            sp.step()
        sims = 0
        wall = 0.0
        for _ in range(args.steps):
            s, _, w = sp.step()
            sims += s
            wall += w
    finally:
        sp.close()
    v = sims / wall
    sample = (f"{sp.workers} worker processes x 1 move ({args.sims} simulations, batch-1 fp32 CPU net, 1 thread each) "
              f"per step, {args.steps} steps, games continue across steps")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{RULES[1]}x{RULES[0]} Connect-{RULES[2]} self-play, {args.sims} simulations/move, policy/value net leaf evaluation",
                   "sims_per_move": args.sims, "board": f"{RULES[1]}x{RULES[0]}", "n_connect": RULES[2], "gravity": RULES[3]},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": sp.workers, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--trees", type=int, default=4096, help="concurrent games per GPU")
    ap.add_argument("--sims", type=int, default=800, help="simulations per move")
    ap.add_argument("--advances", type=int, default=800, help="lock-step advances per step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="bounded CPU baseline sample (0 = skip)")
    ap.add_argument("--unroll", type=int, default=8)
    ap.add_argument("--groups", type=int, default=1, help="tree slices advanced on parallel graph branches")
    ap.add_argument("--max-free", type=int, default=8)
    ap.add_argument("--no-fused", action="store_true", help="three-kernel route instead of az_advance_fused")
    ap.add_argument("--memo-log2", type=int, default=0,
                    help="evaluation memo (the reference's plays_inferences) with 2^n entries; 0 = off (headline)")
    ap.add_argument("--board", default="7x6", help="WxH (headline: 7x6); other boards are extra configurations (C4)")
    ap.add_argument("--connect", type=int, default=4)
    ap.add_argument("--no-gravity", action="store_true")
    ap.add_argument("--game", default="connect_n", choices=["connect_n", "chess"],
                    help="chess = BASELINE config C5 (tools/bench_chess.py); the headline metric is connect_n")
    ap.add_argument("--max-plies", type=int, default=512, help="chess: a game still running after this many plies is a draw")
    args = ap.parse_args()
    if args.game == "chess":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_chess

        return bench_chess.main(args, ClockSampler, load_peaks)
    global RULES
    bw, bh = (int(v) for v in args.board.lower().split("x"))
    RULES = (bw, bh, args.connect, not args.no_gravity)
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # CPU baseline first (rank 0, N = 1 only), before the GPU is busy
    cpu = None
    if world == 1 and args.cpu_seconds > 0:
        from oracle import cpu_baseline

        m = cpu_baseline.measure(RULES, args.sims, "net", seconds=args.cpu_seconds)
        cpu = {"value": m["sims_per_s"], "unit": UNIT, "cores": m["workers"], "kind": "port",
               "sample": (f"{m['workers']} worker processes (os.cpu_count()-1, the reference's own fan-out), each playing "
                          f"{args.sims}-simulation moves of 6x7 Connect-4 with a batch-1 fp32 CPU net for {m['wall_s']:.1f} s "
                          f"({m['moves']} moves, {m['sims']} simulations)")}

    import torch
    import torch.distributed as dist

    from az_b200 import dist as azdist
    from az_b200 import engine, selfplay
    from az_b200.net import PolicyValueNet

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    peaks = load_peaks()
    # cuDNN's autotuner times each candidate once, at the first convolution: bring the clocks to their loaded state first
    # so that it does not pick its algorithms on an idle, boosting GPU (run-to-run spread of the net forward: 0.43-0.48 ms)
    burn = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    for _ in range(200):
        burn @ burn
    torch.cuda.synchronize()
    del burn

    rules = engine.Rules(*RULES)
    T, S, ADV = args.trees, args.sims, args.advances
    torch.manual_seed(0)
    fp32 = PolicyValueNet(rules.height, rules.width, rules.n_actions)
    runner = selfplay.SelfPlayRunner(rules, n_trees=T, sims_per_move=S, net=fp32, games_target=1 << 40,
                                     game_id_base=rank << 40, seed=1234, move_mode="philox", auto_restart=True,
                                     unroll=args.unroll, fin_capacity=4 * T, groups=args.groups, max_free_sims=args.max_free, fused=not args.no_fused,
                                     eval_cache_log2=args.memo_log2)
    flat_dev = runner.net.flat_weights()  # what the trainer rank would broadcast after a training step
    n_w = flat_dev.numel()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step_device():
        if world > 1:
            azdist.broadcast_weights(flat_dev, src=0)
        runner.run(ADV)

    for _ in range(args.warmup):
        step_device()
    runner.fin_clear()
    barrier()
    c0 = runner.totals()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for k in range(args.steps):
        step_device()
        if (k + 1) % 2 == 0:
            runner.fin_clear()  # ring bookkeeping only; samples are consumed in the e2e leg
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    c1 = runner.totals()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    delta = torch.tensor([c1[k] - c0[k] for k in ("sims", "evals", "moves", "games", "depth_sum", "children")],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(delta, op=dist.ReduceOp.SUM)
    ms = float(ms)
    sims, evals, moves, games, depth_sum, children = [float(x) for x in delta]
    runner.check_status()

    # ---- e2e: public API with host buffers (weights up from pinned memory, decoded samples down)
    flat_host = flat_dev.cpu().pin_memory()
    params = list(runner.net.parameters())
    # one untimed pass over the sample path (first use loads the decode kernel and torch's index ops)
    runner.run(args.unroll)
    _warm = runner.finished_device()
    if world > 1:
        _warm = azdist.all_gather_records({k_: v.contiguous() for k_, v in _warm.items()})
    selfplay.decode_samples(rules, _warm)
    runner.fin_clear()
    barrier()
    e0 = runner.totals()
    h2d = d2h = 0
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    dbg = []
    for _ in range(args.steps):
        dbg.append(time.perf_counter())
        flat_dev.copy_(flat_host, non_blocking=True)
        h2d += n_w * 4
        if world > 1:
            azdist.broadcast_weights(flat_dev, src=0)
        off = 0
        with torch.no_grad():
            for p_ in params:
                p_.copy_(flat_dev[off: off + p_.numel()].view_as(p_))
                off += p_.numel()
        runner.run(ADV)
        fin = runner.finished_device()
        if world > 1:
            fin = azdist.all_gather_records({k_: v.contiguous() for k_, v in fin.items()})
        if rank == 0 or world == 1:
            st, po, va = selfplay.decode_samples(rules, fin)
            d2h += st.nbytes + po.nbytes + va.nbytes // 2
        runner.fin_clear()
        dbg.append(time.perf_counter())
    t1.record()
    barrier()
    e1 = runner.totals()
    e2e_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    e2e_sims = torch.tensor([e1["sims"] - e0["sims"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_sims, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_sims) / float(e2e_ms) * 1e3
    if rank == 0 and os.environ.get("AZ_BENCH_DEBUG"):
        print("e2e host wall per step (ms):", [round((dbg[i + 1] - dbg[i]) * 1e3, 1) for i in range(0, len(dbg), 2)],
              "device ms:", float(e2e_ms), file=sys.stderr)

    # ---- roofline of the dominant kernels: the net forward (tensor bound), timed alone with CUDA events
    roof = roof_tree = None
    if rank == 0:
        x = runner.states  # the batch one net call really sees: T / groups positions
        Tg = x.shape[0]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            runner.net(x)
        for _ in range(5):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n_rep = 50
        for _ in range(n_rep):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        net_ms = a.elapsed_time(b) / n_rep
        ach = Tg * runner.flops_per_eval / (net_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_sustained"],
                "traffic": 817.0e6 if (Tg == 4096 and RULES == (7, 6, 4, True)) else None,
                "traffic_note": "DRAM bytes of the 12 tower convolutions per forward at 4096 positions (ncu, steady state, profiles/advance_traffic_r1.json); algorithmic in+out bytes are 1232 MB - the 126 MB L2 keeps part of every activation tensor on chip",
                "kernel": "policy/value net forward timed alone: az_net_stem + 12 cuDNN tcgen05 implicit-GEMM convolutions (cutlass3x_sm100_tensorop, fused bias/ReLU/residual epilogues) + az_net_heads",
                "flops_per_launch": Tg * runner.flops_per_eval, "positions_per_launch": Tg, "ms_per_launch": net_ms, "peak_source": peaks["source"] + ", sustained"}
        # the per-tree kernel alone (HBM bound): algorithmic bytes per tree and launch with the measured mean
        # depth / fan-out.  Fused route: az_advance_fused = heads + tree step + stem (reads the tower output of the
        # tree's leaf, writes the stem output of the next one); else az_step.
        d_bar = depth_sum / max(sims, 1.0)
        k_bar = children / max(evals, 1.0)
        sims_per_tree = sims / max(evals, 1.0)  # simulations finished per evaluated leaf
        A, cells = rules.n_actions, rules.height * rules.width
        tree_bytes = (16 + d_bar * k_bar * 24      # select: root record + per level k children x (16 B record + 8 B prior)
                      + 4 * d_bar + 16 + 4         # path + leaf position + path length written
                      + 4 * d_bar + 16             # path + leaf position re-read at expansion
                      + k_bar * 24                 # expand: k children x 24 B
                      + d_bar * 32                 # backup: 16 B read + 16 B write per path node
                      + 64)                        # per-tree header words
        g0 = runner.groups[0]
        if runner.fused:
            import ctypes

            from az_b200.engine import _ptr, _stream
            from az_b200.native import check, lib

            bytes_per_tree = tree_bytes * sims_per_tree + 2 * cells * 128 * 2  # + tower output in, stem output out (bf16)
            hw = runner.net._heads_arg()

            def launch():
                check(lib().az_advance_fused(g0.engine._h, _ptr(g0.tower_carry), ctypes.byref(hw), _ptr(runner.net.stem_w32),
                                             _ptr(runner.net.stem_b32), _ptr(g0.stem_out), _ptr(g0.valid), _stream()))
            kname = "az::k_advance (az_advance_fused)"
        else:
            bytes_per_tree = tree_bytes * sims_per_tree + cells * 8 + 4 * A + 4  # + bf16 planes out, priors/value in

            def launch():
                g0.engine.step(g0.priors, g0.values, g0.states, g0.valid)
            kname = "az::k_step (az_step)"
        for _ in range(3):
            launch()
        a.record()
        for _ in range(n_rep):
            launch()
        b.record()
        torch.cuda.synchronize()
        step_ms = a.elapsed_time(b) / n_rep
        ach_gbs = Tg * bytes_per_tree / (step_ms * 1e-3) / 1e9
        roof_tree = {"bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ach_gbs / peaks["hbm_gbs"], "traffic": 89.7e6 if runner.fused else None,
                     "traffic_source": "profiles/advance_traffic_r1.json (DRAM bytes of k_advance per launch, ncu, steady state)",
                     "kernel": kname, "bytes_per_tree": bytes_per_tree, "tree_bytes_per_sim": tree_bytes,
                     "mean_depth": d_bar, "mean_children": k_bar, "sims_per_evaluated_leaf": sims_per_tree,
                     "trees_per_launch": Tg, "ms_per_launch": step_ms, "peak_source": peaks["source"]}

    if rank == 0:
        per_adv = args.steps * ADV
        out = {
            "metric": METRIC, "value": sims / ms * 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{'C2' if RULES == (7, 6, 4, True) else 'C4'}: {T} concurrent {rules.height}x{rules.width} Connect-{rules.n} self-play games per GPU x {S} simulations/move, bf16 net leaf evaluation",
                       "games_per_gpu": T, "sims_per_move": S, "advances_per_step": ADV, "groups": args.groups, "max_free_sims": args.max_free, "fused_advance": bool(runner.fused), "evaluation_memo_log2": args.memo_log2, "board": f"{rules.height}x{rules.width}", "n_connect": rules.n, "gravity": rules.gravity,
                       "net": f"4-block 128-filter projection-residual tower, {fp32.n_parameters()} params, random init",
                       "l2": "working set per advance (node pools ~GBs + 177 MB activations per conv) exceeds the 126 MB L2; no flush needed"},
            "leaf_evals_per_sec": evals / ms * 1e3, "memo_hits_per_sec": (c1.get("memo_hits", 0) - c0.get("memo_hits", 0)) / ms * 1e3, "selfplay_moves_per_sec": moves / ms * 1e3,
            "games_finished": games,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // max(args.steps, 1),
                    "d2h_bytes_per_step": d2h // max(args.steps, 1),
                    "what": "per step: weights from pinned host memory -> device (+ NCCL broadcast), 800 advances, finished games decoded to (states f32, policies f64, values) and copied to the host"},
            "gpu_launches": int((runner.launches_per_advance * per_adv + per_adv // args.unroll + args.steps // 2) * world),
            "roofline": roof, "roofline_tree": roof_tree, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
