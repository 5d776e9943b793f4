#!/usr/bin/env python3
"""bench.py - Connect-4 self-play MCTS simulations/s on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's engine
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Workload (config C2 of BASELINE.json, per GPU): 4096 concurrent 6x7 Connect-4 games, 800 simulations
per move, every leaf evaluated by the bf16 policy/value net (random-init weights of the reference
architecture: synthetic), finished games refilled so the batch stays full.  One "step" = 800 lock-step
advances of all trees = one move's worth of simulations for every game (~4.5 M simulations per GPU).  An advance is two
launches of our own kernels: az_step_gather (the tree step) and az_net_forward_trees (the whole net in one tcgen05
kernel, with the idle trees' evaluator-free simulations on two extra warps per CTA).
N > 1: weak scaling by default - every rank owns 4096 games and a net replica (32768 games at N = 8); --scaling strong
--total-games 32768 is config C3 as BASELINE.json words it (32768 games sharded across the GPUs).  The only collectives
are the weight broadcast (side stream, never on the simulation path) and, in the e2e leg, the game-record gather to the
trainer rank.  At N > 1 rank 0 first replays games of other ranks and checks them bit for bit ("multi_rank_parity").
The timed steps follow an untimed pre-roll of more than one game generation (steady state), and the device-resident and
end-to-end steps alternate so that both see the same mix of plies.

Prints ONE JSON line (rank 0).  `value` = whole-job simulations/s, device-timed with everything
resident in HBM; `e2e` = the same metric through the public API with host buffers: every step uploads
the weights from pinned host memory and downloads the decoded training samples of the games that
finished.  `roofline` is the residual tower (tensor bound, the dominant kernel: az_net_tower), timed alone
against the measured BURST peak plus `in_loop_frac` against the sustained one; `roofline_tree` the per-tree kernel
(HBM bound).  `cpu_baseline` = the oracle's Python port of the reference run the way the
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


def multi_rank_parity(dist, azdist, selfplay, rules, fp32, sims, rank, world, n_games=64):
    """SURVEY 4 test 5 on hardware: every rank plays games [rank << 40, +n_games) with its own net replica, the records
    are gathered over NCCL to rank 0, and rank 0 replays the ids of rank 1 and of the last rank locally: moves, visit
    counts, lengths and results must be identical (a game depends on its id, the seed and the weights - not on the rank
    or on what else is in the batch).  Returns the status string printed in the JSON line."""
    import torch

    def play(base):
        r = selfplay.SelfPlayRunner(rules, n_trees=n_games, sims_per_move=sims, net=fp32, games_target=n_games,
                                    game_id_base=base, seed=1234, move_mode="philox", auto_restart=True, unroll=8,
                                    fin_capacity=n_games)
        r.run_until_done(poll_every=512)
        fin = {k: v.clone() for k, v in r.finished_device().items()}
        r.check_status()
        return fin

    mine = play(rank << 40)
    gathered = azdist.gather_records({k: v.contiguous() for k, v in mine.items()}, dst=0)
    status = None
    if rank == 0:
        checked = 0
        for peer in sorted({1, world - 1}):
            want = play(peer << 40)
            sel = (gathered["game_id"] >> 40) == peer
            got = {k: v[sel] for k, v in gathered.items()}
            assert int(sel.sum()) == n_games, "rank %d sent %d games" % (peer, int(sel.sum()))
            og, ow = torch.argsort(got["game_id"]), torch.argsort(want["game_id"])
            for k in ("game_id", "len", "result", "action", "visits", "board"):
                assert torch.equal(got[k][og], want[k][ow]), "multi-rank parity: %s of rank %d differs" % (k, peer)
            checked += n_games
        status = "ok: %d games of ranks %s (%d simulations/move) gathered over NCCL equal their replay on rank 0 (moves, visit counts, results)" % (
            checked, sorted({1, world - 1}), sims)
    return status


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--trees", type=int, default=4096, help="concurrent games per GPU (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong = BASELINE config C3 as written: --total-games sharded across the GPUs")
    ap.add_argument("--total-games", type=int, default=32768, help="concurrent games of the whole job under --scaling strong")
    ap.add_argument("--preroll", type=int, default=48,
                    help="untimed steps before the warm-up: > one game generation (42 plies), so that finished games have "
                         "been refilled and the timed steps see the steady-state mix of plies")
    ap.add_argument("--parity-games", type=int, default=64, help="N > 1: games per rank of the multi-rank parity check (0 = skip)")
    ap.add_argument("--sims", type=int, default=800, help="simulations per move")
    ap.add_argument("--advances", type=int, default=800, help="lock-step advances per step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="bounded CPU baseline sample (0 = skip)")
    ap.add_argument("--unroll", type=int, default=8)
    ap.add_argument("--groups", type=int, default=1, help="tree slices advanced on parallel graph branches")
    ap.add_argument("--max-free", type=int, default=None, help="evaluator-free simulations per tree inside one az_step (default: the runner's choice)")
    ap.add_argument("--extra-sims", type=int, default=0,
                    help="evaluator-free simulations continued beside the net (az_extra_sims on a side stream) for trees without a pending leaf")
    ap.add_argument("--net-tree-sims", type=int, default=None,
                    help="evaluator-free simulations continued INSIDE the net kernel (az_net_forward_trees) for trees without a pending leaf")
    ap.add_argument("--node-capacity", type=int, default=0, help="nodes per tree per pool half (0 = the engine's default sizing)")
    ap.add_argument("--no-fused", action="store_true", help="three-kernel route instead of az_advance_fused")
    ap.add_argument("--no-whole-net", action="store_true",
                    help="az_advance_fused + az_net_tower instead of az_step + az_net_forward (the whole net in one kernel)")
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
    from az_b200.net import PolicyValueNet, flops_per_eval

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    peaks = load_peaks()
    # bring the clocks to their loaded state before anything is autotuned or timed
    burn = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    for _ in range(200):
        burn @ burn
    torch.cuda.synchronize()
    del burn

    rules = engine.Rules(*RULES)
    S, ADV = args.sims, args.advances
    if args.scaling == "strong":
        assert args.total_games % world == 0, "--total-games must divide by the number of GPUs"
        T = args.total_games // world
    else:
        T = args.trees
    torch.manual_seed(0)
    fp32 = PolicyValueNet(rules.height, rules.width, rules.n_actions)

    parity = None
    if world > 1 and args.parity_games > 0:
        parity = multi_rank_parity(dist, azdist, selfplay, rules, fp32, S, rank, world, args.parity_games)
        torch.cuda.empty_cache()

    runner = selfplay.SelfPlayRunner(rules, n_trees=T, sims_per_move=S, net=fp32, games_target=1 << 40,
                                     game_id_base=rank << 40, seed=1234, move_mode="philox", auto_restart=True,
                                     unroll=args.unroll, fin_capacity=4 * T, groups=args.groups, max_free_sims=args.max_free, fused=not args.no_fused, extra_sims=args.extra_sims, net_tree_sims=args.net_tree_sims,
                                     node_capacity=args.node_capacity or None,
                                     eval_cache_log2=args.memo_log2, whole_net=False if args.no_whole_net else None)
    params = list(runner.net.parameters())
    flat_dev = runner.net.flat_weights()            # the live weights as one vector
    flat_trainer = flat_dev.clone()                 # what the trainer rank holds after a training step
    flat_host = flat_dev.cpu().pin_memory()         # ... and on the host (e2e leg)
    n_w = flat_dev.numel()
    bcast = azdist.WeightBroadcaster(n_w, dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def install_weights():
        """flat_dev -> the parameters the captured graphs read (device-to-device, on the compute stream)."""
        off = 0
        with torch.no_grad():
            for p_ in params:
                p_.copy_(flat_dev[off: off + p_.numel()].view_as(p_))
                off += p_.numel()

    def step_device():
        """One step with everything resident: pick up the weights the last broadcast staged, start the next broadcast
        on the side stream (NCCL, N > 1), 800 advances.  The compute stream never waits inside a collective."""
        if world > 1:
            if bcast.take(flat_dev):
                install_weights()
            bcast.start(flat_trainer if rank == 0 else None)
        runner.run(ADV)

    h2d = d2h = 0

    def step_e2e():
        """The same step through the public API with host buffers: weights from pinned host memory (+ broadcast),
        800 advances, the games that finished gathered to the trainer rank, decoded to (states f32, policies f64,
        values) and copied to the host."""
        nonlocal h2d, d2h
        if rank == 0:
            flat_trainer.copy_(flat_host, non_blocking=True)
            h2d += n_w * 4
        if world > 1:
            if bcast.pending:
                bcast.take(flat_dev)  # drop the one the device leg left in flight
            bcast.start(flat_trainer if rank == 0 else None)
            bcast.take(flat_dev)
        else:
            flat_dev.copy_(flat_trainer, non_blocking=True)
        install_weights()
        runner.run(ADV)
        fin = runner.finished_device()
        if world > 1:
            fin = azdist.gather_records({k_: v.contiguous() for k_, v in fin.items()}, dst=0)
        if rank == 0:
            st, po, va = selfplay.decode_samples(rules, fin)
            d2h += st.nbytes + po.nbytes + va.nbytes // 2
        runner.fin_clear()

    # ---- untimed: pre-roll (more than a game generation), then W warm-up steps of both legs
    for k in range(args.preroll):
        step_device()
        if (k + 1) % 2 == 0:
            runner.fin_clear()
    for _ in range(args.warmup):
        step_device()
        step_e2e()
    runner.fin_clear()
    h2d = d2h = 0
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    mk = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    dev_ev = [(mk(), mk()) for _ in range(args.steps)]
    e2e_ev = [(mk(), mk()) for _ in range(args.steps)]
    dev_tot = {k_: 0.0 for k_ in ("sims", "evals", "moves", "games", "depth_sum", "children", "memo_hits")}
    e2e_sims = 0.0
    barrier()
    # ---- timed: K device-resident steps and K end-to-end steps, interleaved step by step so that both legs see the
    # same phase of the games; each step is bracketed by CUDA events on the compute stream
    for k in range(args.steps):
        c0 = runner.totals()
        dev_ev[k][0].record()
        step_device()
        dev_ev[k][1].record()
        c1 = runner.totals()   # reads device counters: synchronises the host with the end of the step
        for k_ in dev_tot:
            dev_tot[k_] += c1.get(k_, 0) - c0.get(k_, 0)
        runner.fin_clear()
        e2e_ev[k][0].record()
        step_e2e()
        e2e_ev[k][1].record()
        c2 = runner.totals()
        e2e_sims += c2["sims"] - c1["sims"]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    runner.check_status()
    my_ms = sum(a.elapsed_time(b) for a, b in dev_ev)
    my_e2e_ms = sum(a.elapsed_time(b) for a, b in e2e_ev)
    ms_max = torch.tensor([my_ms, my_e2e_ms, -my_ms], dtype=torch.float64, device=dev)
    delta = torch.tensor([dev_tot[k_] for k_ in ("sims", "evals", "moves", "games", "depth_sum", "children", "memo_hits")] + [e2e_sims],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(delta, op=dist.ReduceOp.SUM)
    ms, e2e_ms, ms_min = float(ms_max[0]), float(ms_max[1]), -float(ms_max[2])
    sims, evals, moves, games, depth_sum, children, memo_hits, e2e_sims = [float(x) for x in delta]
    e2e_value = e2e_sims / e2e_ms * 1e3

    # ---- roofline of the dominant kernel, timed alone with CUDA events on the launching stream
    roof = roof_tree = None
    if rank == 0:
        x = runner.states  # the batch one net call really sees: T / groups positions
        Tg = x.shape[0]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_rep = 50
        cells = rules.height * rules.width
        tower_flops = 4 * (2 * 2 * cells * 9 * 128 * 128 + 2 * cells * 128 * 128)  # per position: 8 conv3x3 + 4 conv1x1
        g0 = runner.groups[0]
        whole = bool(runner.whole_net)
        if whole:
            # the dominant kernel is the whole net (az_net_forward): leaf planes in, priors / values out
            tower_flops = flops_per_eval(rules.height, rules.width, rules.n_actions)
            pr = torch.empty((Tg, rules.n_actions), dtype=torch.float32, device=dev)
            va = torch.empty(Tg, dtype=torch.float32, device=dev)
            s_in = [torch.randint(0, 2, (Tg, rules.height, rules.width, 4), device=dev).to(torch.bfloat16) for _ in range(4)]
            run_it = lambda i: runner.net(s_in[i % 4], pr, va)  # noqa: E731
        else:
            h_in = [torch.rand((Tg, rules.height, rules.width, 128), device=dev).to(torch.bfloat16) for _ in range(4)]
            run_it = lambda i: runner.net.tower(h_in[i % 4])  # noqa: E731  (inputs rotate through 4 x 44 MB: more than the L2 holds)
        for i in range(8):
            run_it(i)
        a.record()
        for i in range(n_rep):
            run_it(i)
        b.record()
        torch.cuda.synchronize()
        tower_ms = a.elapsed_time(b) / n_rep
        ach = Tg * tower_flops / (tower_ms * 1e-3) / 1e12
        evals_per_s = evals / ms * 1e3
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "tower_traffic_r2.json")
        if os.path.exists(tpath) and Tg == 4096 and RULES == (7, 6, 4, True) and runner.net.fused_tower:
            with open(tpath) as fp:
                tj = json.load(fp)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/tower_traffic_r2.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one k_tower launch at 4096 positions)"
        fused_tower = bool(getattr(runner.net, "fused_tower", False))
        roof = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_burst"],
                "in_loop_frac": evals_per_s / world * tower_flops / 1e12 / peaks["bf16_sustained"],
                "in_loop_note": "leaf evaluations/s per GPU inside the timed steps x tower FLOPs per position / the measured SUSTAINED bf16 peak (the step also contains the per-tree kernel, which runs serially with the tower)",
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": ("az::tower::k_tower<true> (az_net_forward): the whole policy/value net - stem, 4-block residual tower, both heads - in one persistent tcgen05 kernel, activations resident in shared memory / TMEM" if whole
                           else "az::tower::k_tower<false> (az_net_tower): the whole 4-block residual tower, one persistent tcgen05 kernel, activations resident in shared memory / TMEM" if fused_tower
                           else "12 cuDNN tcgen05 implicit-GEMM convolutions (cutlass3x_sm100_tensorop) of the residual tower"),
                "flops_per_launch": Tg * tower_flops, "positions_per_launch": Tg, "ms_per_launch": tower_ms,
                "share_of_net_flops": tower_flops / flops_per_eval(rules.height, rules.width, rules.n_actions),
                "peak_source": peaks["source"] + ", burst (kernel timed alone)"}
        # the per-tree kernel alone (HBM bound): algorithmic bytes per tree and launch with the measured mean
        # depth / fan-out.  Fused route: az_advance_fused = heads + tree step + stem (reads the tower output of the
        # tree's leaf, writes the stem output of the next one); else az_step.
        d_bar = depth_sum / max(sims, 1.0)
        k_bar = children / max(evals, 1.0)
        sims_per_tree = sims / max(evals, 1.0)  # simulations finished per evaluated leaf
        A = rules.n_actions
        tree_bytes = (16 + d_bar * k_bar * 24      # select: root record + per level k children x (16 B record + 8 B prior)
                      + 4 * d_bar + 16 + 4         # path + leaf position + path length written
                      + 4 * d_bar + 16             # path + leaf position re-read at expansion
                      + k_bar * 24                 # expand: k children x 24 B
                      + d_bar * 32                 # backup: 16 B read + 16 B write per path node
                      + 64)                        # per-tree header words
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
                     "frac": ach_gbs / peaks["hbm_gbs"], "traffic": None,
                     "kernel": kname, "bytes_per_tree": bytes_per_tree, "tree_bytes_per_sim": tree_bytes,
                     "mean_depth": d_bar, "mean_children": k_bar, "sims_per_evaluated_leaf": sims_per_tree,
                     "trees_per_launch": Tg, "ms_per_launch": step_ms, "peak_source": peaks["source"]}

    if rank == 0:
        per_adv = args.steps * ADV
        games_job = T * world
        cfg_name = "C3" if args.scaling == "strong" else ("C2" if RULES == (7, 6, 4, True) else "C4")
        out = {
            "metric": METRIC, "value": sims / ms * 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{cfg_name}: {games_job} concurrent {rules.height}x{rules.width} Connect-{rules.n} self-play games ({T} per GPU) x {S} simulations/move, bf16 net leaf evaluation",
                       "games_per_gpu": T, "games_total": games_job, "sims_per_move": S, "advances_per_step": ADV, "groups": args.groups, "max_free_sims": int(runner.max_free_sims), "extra_sims_beside_net": args.extra_sims, "tree_sims_inside_net": int(runner.net_tree_sims), "node_capacity": int(runner.engine.cfg.node_capacity), "fused_advance": bool(runner.fused), "whole_net_kernel": bool(runner.whole_net),
                       "fused_tower": bool(getattr(runner.net, "fused_tower", False)), "evaluation_memo_log2": args.memo_log2, "board": f"{rules.height}x{rules.width}", "n_connect": rules.n, "gravity": rules.gravity,
                       "net": f"4-block 128-filter projection-residual tower, {fp32.n_parameters()} params, random init",
                       "preroll_steps": args.preroll,
                       "phase": "steady state: the timed steps follow an untimed pre-roll of more than one game generation; the device-resident and the end-to-end steps alternate, so both see the same mix of plies",
                       "l2": "working set per advance (node pools ~GBs + 88 MB of activations in and out of the tower) exceeds the 126 MB L2; no flush needed"},
            "leaf_evals_per_sec": evals / ms * 1e3, "memo_hits_per_sec": memo_hits / ms * 1e3, "selfplay_moves_per_sec": moves / ms * 1e3,
            "games_finished": games,
            "rank_ms_per_step": {"min": ms_min / args.steps, "max": ms / args.steps},
            "multi_rank_parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // max(args.steps, 1),
                    "d2h_bytes_per_step": d2h // max(args.steps, 1),
                    "what": "per step: weights from pinned host memory -> device (+ NCCL broadcast on a side stream), 800 advances, finished games gathered to the trainer rank, decoded to (states f32, policies f64, values) and copied to the host"},
            "gpu_launches": int((2 * runner.launches_per_advance * per_adv + args.steps) * world),
            "roofline": roof, "roofline_tree": roof_tree, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
