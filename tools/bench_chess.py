#!/usr/bin/env python3
"""Chess self-play measurement (BASELINE config C5): `python bench.py --game chess ...` lands here.

Workload per GPU: `--trees` concurrent chess games x `--sims` simulations per move, every leaf evaluated by the bf16
policy/value net (the reference's architecture at 8x8x118 -> 1 880 actions, random init: synthetic), finished games
refilled.  One step = `--advances` lock-step advances.  Prints ONE JSON line with the same keys as the Connect-4
bench: `value` (device resident), `e2e` (weights up from pinned host memory, finished plies decoded to training
arrays and copied to the host), `roofline` (net forward, tensor bound), `roofline_tree` (az_chess_step alone, HBM
bound), `movegen` (az_chess_perft: legal-move generation + make + mirror, positions/s) and `cpu_baseline` (the C
oracle's MCTS over the mailbox rules on one host core - WITHOUT a net, so it overstates what the reference's CPU
path with python-chess and a batch-1 net would do).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "custom-alphazero_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "chess_selfplay_mcts_simulations_per_sec"
UNIT = "sims/s"


def cpu_port_baseline(sims, seconds):
    """oracle/c/chess_oracle.c MCTS (hash evaluator, no net), one thread, for about `seconds`."""
    from oracle import chess_ref as cr

    t0 = time.perf_counter()
    total = 0
    plies = 0
    games = 0
    while time.perf_counter() - t0 < seconds:
        g = cr.mcts_game(sims=sims, evaluator="hash", max_plies=24)
        total += g["sims"]
        plies += g["plies"]
        games += 1
    wall = time.perf_counter() - t0
    return {"value": total / wall, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": (f"C oracle MCTS over the mailbox rules, hash evaluator instead of a net, 1 thread, {games} games cut at 24 "
                       f"plies x {sims} simulations/move for {wall:.1f} s ({total} simulations); the reference's own chess path "
                       "needs python-chess (absent) and fails inside MCTS (SURVEY.md section 2 #15)")}


def main(args, ClockSampler, load_peaks):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            m = cpu_port_baseline(args.sims, max(args.cpu_seconds, 5.0))
            print(json.dumps({"impl": "reference", "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": args.gpus,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
                              "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                              "config": {"workload": f"chess self-play, {args.sims} simulations/move"},
                              "cpu_baseline": m, "e2e": {"value": m["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                                         "d2h_bytes_per_step": 0}}))
        return
    cpu = cpu_port_baseline(args.sims, args.cpu_seconds) if (world == 1 and args.cpu_seconds > 0) else None

    import torch
    import torch.distributed as dist

    from az_b200 import chess
    from az_b200 import dist as azdist
    from az_b200.chess_selfplay import ChessSelfPlayRunner, chess_net

    assert torch.cuda.is_available(), "bench needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    peaks = load_peaks()
    burn = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)  # loaded clocks before cuDNN's autotuner times anything
    for _ in range(200):
        burn @ burn
    torch.cuda.synchronize()
    del burn
    T, S, ADV = args.trees, args.sims, args.advances
    torch.manual_seed(0)
    fp32 = chess_net()
    runner = ChessSelfPlayRunner(n_trees=T, sims_per_move=S, net=fp32, games_target=1 << 40, game_id_base=rank << 40,
                                 seed=1234, move_mode="philox", auto_restart=True, unroll=args.unroll,
                                 max_free_sims=args.max_free or 8, max_plies=args.max_plies,
                                 sample_capacity=max(1024, T * (2 + 2 * ADV // S)))
    flat_dev = runner.net.flat_weights()
    n_w = flat_dev.numel()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step_device():
        if world > 1:
            azdist.broadcast_weights(flat_dev, src=0)
        runner.run(ADV)
        runner.engine.rings_clear()  # ring bookkeeping only; samples are consumed in the e2e leg

    for _ in range(args.warmup):
        step_device()
    barrier()
    c0 = runner.totals()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    c1 = runner.totals()
    runner.engine.check_status()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    keys = ("sims", "evals", "moves", "games", "depth_sum", "children")
    delta = torch.tensor([c1[k] - c0[k] for k in keys], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(delta, op=dist.ReduceOp.SUM)
    ms = float(ms)
    sims, evals, moves, games, depth_sum, children = [float(x) for x in delta]

    # ---- e2e: weights from pinned host memory, finished plies decoded and downloaded
    flat_host = flat_dev.cpu().pin_memory()
    params = list(runner.net.parameters())
    runner.run(args.unroll)
    runner.collect()
    barrier()
    e0 = runner.totals()
    h2d = d2h = 0
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
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
        if world > 1:  # the replay-buffer gather: every rank's finished plies reach the trainer rank, which decodes them
            got = runner.collect_all_ranks()
        else:
            got = runner.collect()
        if got is not None:
            d2h += got[0].nbytes + got[1].nbytes + got[2].nbytes
    t1.record()
    barrier()
    e1 = runner.totals()
    e2e_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    e2e_sims = torch.tensor([e1["sims"] - e0["sims"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_sims, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_sims) / float(e2e_ms) * 1e3

    roof = roof_tree = movegen = None
    if rank == 0:
        n_rep = 30
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if runner.stem_from_boards:
                runner.net.forward_from_stem(runner.net.chess_stem(runner.engine.view("leaf_pos")), runner.priors, runner.values)
            else:
                runner.net(runner.states, runner.priors, runner.values)
        for _ in range(5):
            g.replay()
        a.record()
        for _ in range(n_rep):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        net_ms = a.elapsed_time(b) / n_rep
        ach = T * runner.flops_per_eval_executed / (net_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_sustained"], "traffic": None,
                "kernel": "policy/value net forward timed alone: az_chess_stem_tc (from the 64-byte boards), "
                          + ("az_net_tower (the 4-block residual tower as one persistent tcgen05 kernel, two 8x8 positions per tile)"
                             if getattr(runner.net, "fused_tower", False) else
                             "12 cuDNN tcgen05 implicit-GEMM convolutions with fused epilogues")
                          + ", az_net_head_convs, az_net_dense_heads (tcgen05: policy over 1 880 actions + softmax, value MLP)",
                "flops_per_launch": T * runner.flops_per_eval_executed, "flops_per_launch_reference_net": T * runner.flops_per_eval,
                "flops_note": "executed FLOPs: with tail_planes the stem multiplies the 34 planes that can be non-zero on the self-play path instead of 118",
                "positions_per_launch": T, "ms_per_launch": net_ms,
                "peak_source": peaks["source"] + ", sustained"}
        # az_chess_step alone: algorithmic bytes per tree and launch from the measured mean depth / fan-out
        d_bar = depth_sum / max(sims, 1.0)
        k_bar = children / max(evals, 1.0)
        sims_per_leaf = sims / max(evals, 1.0)
        per_sim = (16 + d_bar * k_bar * 24 + d_bar * 2   # select: k children x (16 B record + 8 B prior), one action per level
                   + 8 * d_bar + 2 * (64 + 256)          # stored path, leaf position and legal mask out and back
                   + k_bar * 26                          # expand: k children x (record + prior + action)
                   + d_bar * 32 + 64 + 64)               # backup RMW, root position, header words
        bytes_per_tree = per_sim * sims_per_leaf + 4 * k_bar + 4 + (0 if runner.stem_from_boards else 2 * 64 * runner.states.shape[-1])  # + legal priors / value in (+ bf16 planes out)

        def launch():
            runner.engine.step(runner.priors, runner.values, None if runner.stem_from_boards else runner.states, runner.valid,
                               runner.plane_first)
        for _ in range(3):
            launch()
        a.record()
        for _ in range(n_rep):
            launch()
        b.record()
        torch.cuda.synchronize()
        step_ms = a.elapsed_time(b) / n_rep
        ach_gbs = T * bytes_per_tree / (step_ms * 1e-3) / 1e9
        roof_tree = {"bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ach_gbs / peaks["hbm_gbs"], "traffic": None, "kernel": "azc::k_chess_step (az_chess_step)",
                     "bytes_per_tree": bytes_per_tree, "tree_bytes_per_sim": per_sim, "mean_depth": d_bar,
                     "mean_children": k_bar, "sims_per_evaluated_leaf": sims_per_leaf, "trees_per_launch": T,
                     "ms_per_launch": step_ms, "peak_source": peaks["source"]}
        # move generator alone: depth-3 move paths from the 197 281 positions four plies into the game (one thread each)
        front = torch.from_numpy(chess.position_from_fen()[None].view("int64")).to(dev)
        for _ in range(4):
            mask, _, _ = chess.chess_legal(front)
            idx, act = torch.nonzero(mask, as_tuple=True)
            front, _ = chess.chess_play(front[idx], act.to(torch.int32))
        chess.chess_perft(front, 2)
        torch.cuda.synchronize()
        a.record()
        nodes = chess.chess_perft(front, 3)
        b.record()
        torch.cuda.synchronize()
        pf_ms = a.elapsed_time(b)
        total = int(nodes.sum())
        movegen = {"kernel": "azc::k_chess_perft (az_chess_perft)", "roots": int(front.shape[0]), "depth": 3,
                   "move_paths": total, "expected_startpos_perft7": 3195901860, "ms": pf_ms,
                   "leaf_positions_per_sec": total / (pf_ms * 1e-3)}

    if rank == 0:
        per_adv = args.steps * ADV
        out = {
            "metric": METRIC, "value": sims / ms * 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"C5: {T} concurrent chess self-play games per GPU x {S} simulations/move, bf16 net leaf evaluation",
                       "games_per_gpu": T, "sims_per_move": S, "advances_per_step": ADV, "max_free_sims": args.max_free or 8,
                       "max_plies": args.max_plies, "leaf_planes": int(runner.states.shape[-1]),
                       "net": f"4-block 128-filter projection-residual tower, 8x8x118 in, 1880 actions, {fp32.n_parameters()} params, random init",
                       "l2": "working set per advance (node pools ~GBs + 134 MB activations per conv at 8192 trees) exceeds the 126 MB L2"},
            "leaf_evals_per_sec": evals / ms * 1e3, "selfplay_moves_per_sec": moves / ms * 1e3, "games_finished": games,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // max(args.steps, 1),
                    "d2h_bytes_per_step": d2h // max(args.steps, 1),
                    "what": "per step: weights from pinned host memory -> device (+ NCCL broadcast), the advances, every finished ply decoded to (state f32 [8,8,118], policy f64 [1880], value) and copied to the host"},
            "gpu_launches": int(runner.launches_per_advance * per_adv * world),
            "roofline": roof, "roofline_tree": roof_tree, "movegen": movegen, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
