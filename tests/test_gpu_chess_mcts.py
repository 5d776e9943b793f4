"""GPU parity tests of the chess search engine (az_chess_search / az_chess_step / az_chess_move) against the C oracle's
MCTS over the mailbox rules (oracle/c/chess_oracle.c: co_mcts_game): per ply the legal actions, root visit counts
and chosen move must be identical, for whole games, with the in-kernel evaluators and through the external route."""
import numpy as np
import pytest

from oracle import chess_ref as cr

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

MASK = (1 << 64) - 1


def _play_games(eng, max_iters, route="search", evaluator=None):
    """Runs every tree to the end of its game; returns {game: [(k, act, n, choice)...]}, {game: (len, result)}."""
    from az_b200 import native

    per_game, fin = {}, {}
    T = eng.n_trees
    states = torch.full((T, 8, 8, 120), 7.0, dtype=torch.bfloat16, device=eng.device)  # 2 pad planes: must come back as zeros
    valid = torch.zeros(T, dtype=torch.int32, device=eng.device)
    for _ in range(max_iters):
        ph = eng.phases()
        if bool((ph == native.AZ_PHASE_IDLE).all()):
            break
        if route == "search":
            eng.search()
        else:
            pri = val = None
            for _adv in range(100000):
                eng.step(pri, val, states, valid)
                if bool((eng.phases() != native.AZ_PHASE_SEARCH).all()):
                    break
                pri, val = evaluator(eng, states, valid) if bool(valid.any()) else (None, None)
        eng.check_status()
        eng.move()
        eng.check_status()
        d = eng.drain()
        for i in range(len(d["k"])):
            per_game.setdefault(int(d["game"][i]), []).append(
                (int(d["ply"][i]), int(d["k"][i]), d["act"][i].copy(), d["n"][i].copy(), int(d["choice"][i])))
        for g, ln, r in zip(d["fin_game"], d["fin_len"], d["fin_result"]):
            fin[int(g)] = (int(ln), int(r))
    return per_game, fin


def _assert_game_equals(plies, fin, want):
    assert fin == (want["plies"], want["result"])
    assert [p[0] for p in plies] == list(range(want["plies"]))
    for ply, k, act, n, choice in plies:
        assert k == int(want["k"][ply]), ply
        assert np.array_equal(act[:k], want["act"][ply][:k]), ply
        assert np.array_equal(n[:k], want["n"][ply][:k]), (ply, n[:k], want["n"][ply][:k])
        assert choice == int(want["choice"][ply]), ply


@pytest.mark.parametrize("evaluator,prior_mode,sims,max_plies", [
    ("hash", "f64", 60, 60), ("uniform", "f64", 100, 16), ("hash", "f32", 40, 300), ("uniform", "f32", 64, 10)])
def test_in_kernel_search_reproduces_the_oracle_game(evaluator, prior_mode, sims, max_plies):
    from az_b200.chess_engine import ChessTreeEngine

    want = cr.mcts_game(sims=sims, evaluator=evaluator, prior_mode=prior_mode, max_plies=max_plies)
    eng = ChessTreeEngine(n_trees=5, sims_per_move=sims, eval_mode=evaluator, prior_mode=prior_mode, max_plies=max_plies)
    per_game, fin = _play_games(eng, max_plies + 2)
    assert sorted(per_game) == list(range(5))
    for g in range(5):
        _assert_game_equals(per_game[g], fin[g], want)
    tot = eng.totals()
    assert tot["sims"] == 5 * want["sims"] and tot["evals"] == 5 * want["evals"] and tot["games"] == 5


def test_sampled_moves_and_compaction():
    """np.random.choice semantics with host-supplied draws, a different game in every tree; pools so small that every
    re-root takes the compaction path must give the same games as roomy pools."""
    from az_b200.chess_engine import ChessTreeEngine

    T, sims, P = 6, 48, 80
    rs = np.random.RandomState(5)
    u = rs.random_sample((T, P))
    wants = [cr.mcts_game(sims=sims, evaluator="hash", max_plies=P, uniforms=u[t], greedy_idx=30) for t in range(T)]
    assert len({tuple(w["choice"][:6]) for w in wants}) > 1
    for cap in (None, 8192):
        eng = ChessTreeEngine(n_trees=T, sims_per_move=sims, eval_mode="hash", prior_mode="f64", move_mode="host_uniforms",
                              max_plies=P, index_move_greedy=30, node_capacity=cap)
        eng.set_uniforms(u)
        per_game, fin = _play_games(eng, P + 2)
        for t in range(T):
            _assert_game_equals(per_game[t], fin[t], wants[t])
        if cap is not None:
            assert eng.totals()["reroot_nodes"] > 0


def _host_hash_evaluator(eng, states, valid):
    """The hash evaluator computed on the host from the leaf positions the engine stored (float64 priors)."""
    from az_b200 import chess

    pos = eng.view("leaf_pos").cpu().numpy().view(np.uint64)
    T = eng.n_trees
    pri = np.zeros((T, 1880), dtype=np.float64)
    val = np.zeros(T, dtype=np.float64)
    v = valid.cpu().numpy()
    # the planes handed to the evaluator are those of the stored leaf
    idx = np.nonzero(v)[0]
    if len(idx):
        enc = chess.chess_encode(pos[idx])
        got = states[torch.as_tensor(idx, device=states.device)].float().cpu().numpy()
        assert np.array_equal(got[..., :118], enc) and not got[..., 118:].any()
    a = np.arange(1880, dtype=np.uint64)
    for t in idx:
        s = cr.from_pos(pos[t])
        h = 0xCBF29CE484222325
        for sq in range(64):
            h = ((h ^ (s.sq[sq] + 7)) * 0x100000001B3) & MASK
        h = ((h ^ ((s.castling & 15) | ((s.ep + 1) << 4))) * 0x100000001B3) & MASK
        with np.errstate(over="ignore"):
            m = (np.uint64(h) ^ (a * np.uint64(0x9E3779B97F4A7C15))) * np.uint64(0xFF51AFD7ED558CCD)
        pri[t] = ((m >> np.uint64(40)) % np.uint64(1000) + np.uint64(1)).astype(np.float64)
        val[t] = (float((h >> 20) % 2001) - 1000.0) / 1000.0
    return (torch.as_tensor(pri, device=states.device), torch.as_tensor(val, device=states.device))


def test_external_route_equals_the_oracle():
    from az_b200.chess_engine import ChessTreeEngine

    sims, P = 30, 24
    want = cr.mcts_game(sims=sims, evaluator="hash", prior_mode="f64", max_plies=P)
    eng = ChessTreeEngine(n_trees=3, sims_per_move=sims, eval_mode="external", max_plies=P, max_free_sims=2)
    per_game, fin = _play_games(eng, P + 2, route="step", evaluator=_host_hash_evaluator)
    for g in range(3):
        _assert_game_equals(per_game[g], fin[g], want)


def test_refill_and_decoded_samples():
    """auto_restart keeps the batch full; decoded policies / values follow self_play.py:63-78."""
    from az_b200 import chess
    from az_b200.chess_engine import ChessTreeEngine, decode_samples, sample_values

    sims, P, G = 24, 12, 7
    eng = ChessTreeEngine(n_trees=3, sims_per_move=sims, eval_mode="hash", prior_mode="f64", max_plies=P, games_target=G,
                          auto_restart=True, index_move_greedy=4)
    want = cr.mcts_game(sims=sims, evaluator="hash", max_plies=P, greedy_idx=4)
    per_game, fin = _play_games(eng, 10 * P)
    assert sorted(fin) == list(range(G)) and sorted(per_game) == list(range(G))
    for g in range(G):
        _assert_game_equals(per_game[g], fin[g], want)
    # one game's samples through the decoder
    plies = per_game[0]
    d = {"pos": None, "k": np.array([p[1] for p in plies], dtype=np.int32),
         "act": np.stack([p[2] for p in plies]), "n": np.stack([p[3] for p in plies]),
         "choice": np.array([p[4] for p in plies], dtype=np.int32), "game": np.zeros(len(plies), dtype=np.int64),
         "ply": np.arange(len(plies), dtype=np.int32), "fin_game": np.array([0]), "fin_len": np.array([fin[0][0]]),
         "fin_result": np.array([fin[0][1]])}
    # parent positions: replay the chosen moves from the start position with the environment kernel
    pos = [chess.position_from_fen()]
    for p in plies[:-1]:
        nxt, st = chess.chess_play(pos[-1][None], np.array([p[4] & 0xFFFF], dtype=np.int32))
        pos.append(nxt[0])
    d["pos"] = np.stack(pos)
    states, policies = decode_samples(d)
    assert np.array_equal(states.cpu().numpy(), chess.chess_encode(d["pos"]))
    pol = policies.cpu().numpy()
    for i, (ply, k, act, n, choice) in enumerate(plies):
        ref = np.zeros(1880)
        if choice >> 16:
            ref[act[int(np.argmax(n[:k]))]] = 1.0
        else:
            ref[act[:k].astype(np.int64)] = n[:k] / n[:k].sum()
        assert np.array_equal(pol[i], ref), i
    vals, known = sample_values(d)
    L, r = fin[0]
    assert known.all() and list(vals) == [r * (1 if (L - 1 - i) % 2 == 0 else -1) for i in range(L)]


def test_chess_net_bf16_against_fp32():
    """The chess-shaped net (8x8x118 -> 1 880 actions): bf16 GPU path (cuDNN tower) against the fp32 module."""
    from az_b200.chess_selfplay import chess_net
    from az_b200.net import InferenceNet, randomise_bn

    torch.manual_seed(1)
    net = randomise_bn(chess_net()).eval()
    x = (torch.rand(64, 8, 8, 118) < 0.1).float()
    with torch.no_grad():
        p32, v32 = net(x)
    inf = InferenceNet(net, dtype=torch.bfloat16, device="cuda")
    p16, v16 = inf(x.cuda().to(torch.bfloat16))
    dp = (p16.cpu() - p32).abs().max().item()
    dv = (v16.cpu() - v32.reshape(-1)).abs().max().item()
    assert p16.shape == (64, 1880) and abs(float(p16.sum()) - 64.0) < 1e-2
    assert dp < 2e-2 and dv < 8e-2, (dp, dv)  # bf16 activations through 13 convolutions; tolerance as tests/test_gpu_net.py


def test_chess_selfplay_runner_with_the_net():
    """Tiny end-to-end: 16 trees, 12 simulations per move, games cut at 10 plies, bf16 net, CUDA graph."""
    from az_b200.chess_selfplay import ChessSelfPlayRunner

    torch.manual_seed(0)
    r = ChessSelfPlayRunner(n_trees=16, sims_per_move=12, games_target=24, max_plies=10, unroll=4)
    states, policies, values, known = r.run_until_done(poll_every=32, max_advances=20000)
    tot = r.totals()
    assert tot["games"] == 24 and tot["moves"] == 240 and states.shape == (240, 8, 8, 118) and policies.shape == (240, 1880)
    assert known.all() and np.allclose(policies.sum(-1), 1.0) and set(np.unique(values)).issubset({-1, 0, 1})
    assert tot["sims"] >= 240 * 12
    # planes of a first ply: Board() itself, seven empty history entries (chess/board.py:37-40); every later ply carries the
    # initial position in entry 6 (planes 84-97)
    empty6 = np.abs(states[:, :, :, 84:98]).sum((1, 2, 3)) == 0
    assert int(empty6.sum()) == 24
    start = cr.full_state(cr.start_state(), [None] * 7)
    assert all(np.array_equal(x.astype(np.float64), start) for x in states[empty6])
    later = states[~empty6]
    assert all(np.array_equal(x[:, :, 84:98].astype(np.float64), start[:, :, 98:112]) for x in later[:50])


def test_sharding_invariance():
    """A game's moves depend on (seed, game id, ply) only: trees hosted by another engine / rank replay the same games."""
    from az_b200.chess_engine import ChessTreeEngine

    kw = dict(sims_per_move=24, eval_mode="hash", prior_mode="f64", move_mode="philox", max_plies=14, seed=99,
              index_move_greedy=30)
    whole, fin_w = _play_games(ChessTreeEngine(n_trees=4, **kw), 16)
    part, fin_p = _play_games(ChessTreeEngine(n_trees=2, game_id_base=2, **kw), 16)
    assert sorted(part) == [2, 3]
    for g in (2, 3):
        assert fin_w[g] == fin_p[g]
        assert [(p[0], p[1], p[4]) for p in whole[g]] == [(p[0], p[1], p[4]) for p in part[g]]
        assert all(np.array_equal(a[3], b[3]) for a, b in zip(whole[g], part[g]))
    assert [p[4] for p in whole[0]] != [p[4] for p in whole[1]]  # different games really differ


def test_chess_closed_loop_trains():
    """Self-play -> sample ring -> replay window -> SGD step (the reference's losses) -> new weights in the runner."""
    from az_b200.chess_selfplay import ChessSelfPlayRunner, chess_net, chess_training_loop
    from az_b200.train import ReplayWindow, Trainer

    torch.manual_seed(0)
    net = chess_net()
    runner = ChessSelfPlayRunner(n_trees=16, sims_per_move=8, net=net, games_target=16, max_plies=8, unroll=4,
                                 auto_restart=False)
    trainer = Trainer(net)
    window = ReplayWindow(8, 8, 1880, capacity=512, planes=118)
    before = runner.net.flat_weights().clone()
    import az_b200.train as T

    old_min, old_batch = T.MIN_TRAINING_SIZE, T.BATCH_SIZE
    T.MIN_TRAINING_SIZE = 64
    try:
        hist = chess_training_loop(runner, trainer, window, iterations=2, exclude_null_games=False,
                                   rng=np.random.RandomState(0), max_advances=4000)
    finally:
        T.MIN_TRAINING_SIZE, T.BATCH_SIZE = old_min, old_batch
    assert len(window) == 256 and len(hist) >= 1 and all(np.isfinite(h["loss"]) for h in hist)
    assert hist[0]["policy_loss"] > 1.0  # cross-entropy against visit distributions over ~20 legal moves
    assert not torch.equal(before, runner.net.flat_weights())


def test_entry_point_plays_chess(tmp_path, monkeypatch):
    """`python -m custom_alphazero.self_play` with ConfigGeneral.game = "chess": one stand-alone iteration writes
    samples.npz with the reference's keys and the chess shapes."""
    import importlib
    import sys

    from custom_alphazero import config

    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(config.ConfigGeneral, "game", "chess")
    monkeypatch.setattr(config.ConfigSelfPlay, "mcts_iterations", 8)
    monkeypatch.setattr(config.ConfigSelfPlay, "exclude_null_games", False)
    monkeypatch.setattr(config.ConfigB200, "concurrent_games", 8)
    monkeypatch.setattr(config.ConfigB200, "games_per_iteration", 12)
    monkeypatch.setattr(config.ConfigB200, "chess_max_plies", 6)
    monkeypatch.setattr(config.ConfigB200, "graph_unroll", 2)
    monkeypatch.setattr(config.ConfigServing, "serving_address", "http://127.0.0.1:9")  # nothing listens: stand-alone
    for name in ("custom_alphazero.self_play", "custom_alphazero.mcts.mcts", "custom_alphazero.mcts.chess_mcts"):
        sys.modules.pop(name, None)
    sp = importlib.import_module("custom_alphazero.self_play")
    try:
        assert sp.Board.__module__ == "custom_alphazero.chess.board" and len(sp.get_all_possible_moves()) == 1880
        sp.main(max_iterations=1)
        files = list(tmp_path.rglob("samples.npz"))
        assert len(files) == 1 and "chess" in str(files[0])
        data = np.load(files[0])
        assert set(data.files) == {"states", "policies", "values"}
        assert data["states"].shape == (72, 8, 8, 118) and data["policies"].shape == (72, 1880) and len(data["values"]) == 72
        # play_game: one game through the drop-in search object (a host round trip per simulation)
        monkeypatch.setattr(config.ConfigB200, "chess_max_plies", 512)
        cut = {"n": 0}
        real_over = sp.Board.is_game_over

        def over_after_four(self):  # stop the plumbing check after four plies
            cut["n"] += 1
            return cut["n"] > 4 or real_over(self)

        monkeypatch.setattr(sp.Board, "is_game_over", over_after_four)
        monkeypatch.setattr(sp.Board, "get_result", lambda self, keep_same_player=False: 0)
        states, policies, rewards, search = sp.play_game(0, sp.get_all_possible_moves(), 8, "x")
        assert states.shape == (4, 8, 8, 118) and policies.shape == (4, 1880) and len(rewards) == 4 and search.model is None
        assert np.allclose(policies.sum(-1), 1.0)
    finally:
        for name in ("custom_alphazero.self_play", "custom_alphazero.mcts.mcts", "custom_alphazero.mcts.chess_mcts"):
            sys.modules.pop(name, None)  # the next importer gets the Connect-N flavour again


def test_fused_dense_heads_against_float64():
    """az_net_dense_heads (tcgen05: policy GEMM + softmax, value MLP + tanh) against the same arithmetic in float64 on
    the operands the kernel sees (bf16-rounded features and weights, float32 biases)."""
    import ctypes

    from az_b200 import native

    torch.manual_seed(2)
    A = 1880
    lib = native.lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    wp = (torch.randn(A, 128, device="cuda") * 0.5).to(torch.bfloat16)
    wp_pad = torch.cat([wp, torch.zeros(1920 - A, 128, device="cuda", dtype=torch.bfloat16)]).contiguous()
    bp = torch.rand(A, device="cuda") * 2 - 1
    w1 = (torch.randn(256, 64, device="cuda") * 0.2).to(torch.bfloat16).contiguous()
    b1 = torch.rand(256, device="cuda") - 0.5
    w2 = torch.randn(256, device="cuda") * 0.2
    b2 = torch.tensor([0.1], device="cuda")
    for n in (4096, 1, 127, 129, 300):
        hd = torch.relu(torch.randn(n, 64, 3, device="cuda")).contiguous()
        priors = torch.full((n + 1, A), 7.0, device="cuda")
        values = torch.full((n + 1,), 7.0, device="cuda")
        scratch = torch.empty((n, native.AZ_DENSE_HEAD_SPLITS, 2), device="cuda")
        native.check(lib.az_net_dense_heads(P(hd), P(wp_pad), P(bp), P(w1), P(b1), P(w2), P(b2), n, 64, A, P(priors), P(values),
                                            P(scratch), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        f = hd.to(torch.bfloat16).double()
        pf = f[:, :, :2].reshape(n, 128)            # Keras Flatten on [cells][2 planes]
        vf = f[:, :, 2]
        ref_p = torch.softmax(pf @ wp.double().t() + bp.double(), dim=-1)
        ref_v = torch.tanh(torch.relu(vf @ w1.double().t() + b1.double()) @ w2.double() + b2.double())
        dp = (priors[:n].double() - ref_p).abs().max().item()
        dv = (values[:n].double() - ref_v).abs().max().item()
        assert dp <= 2e-6 + 2e-5 * ref_p.max().item(), (n, dp, ref_p.max().item())
        assert dv <= 2e-5, (n, dv)
        assert bool((priors[n] == 7.0).all()) and float(values[n]) == 7.0  # nothing written past the last row
        assert float(ref_p.max()) > 50.0 / A  # sharp rows: the softmax statistics really matter


def test_rings_and_pools_fail_loudly_and_recover():
    """A full sample ring leaves trees READY until the host drains it; a full finished-game ring parks the game
    (STALLED) and delivers it afterwards; an exhausted node pool sets the sticky flag instead of corrupting the tree."""
    from az_b200 import native
    from az_b200.chess_engine import ChessTreeEngine
    from az_b200.native import NativeError

    want = cr.mcts_game(sims=16, evaluator="hash", max_plies=6)
    for sample_cap, fin_cap, expect in ((4, 8, native.AZ_PHASE_READY), (8, 2, native.AZ_PHASE_STALLED)):
        eng = ChessTreeEngine(n_trees=6, sims_per_move=16, eval_mode="hash", prior_mode="f64", max_plies=6,
                              sample_capacity=sample_cap, fin_capacity=fin_cap)
        plies, fin = {}, {}
        seen = False
        for _ in range(200):
            if bool((eng.phases() == native.AZ_PHASE_IDLE).all()):
                break
            eng.search()
            eng.move()
            seen |= bool((eng.phases() == expect).any())  # READY: no ring slot, the move waits; STALLED: game over, ring full
            assert int(eng.view("smp_count")[0]) <= sample_cap and int(eng.view("fin_count")[0]) <= fin_cap
            d = eng.drain()
            for i in range(len(d["k"])):
                plies.setdefault(int(d["game"][i]), []).append((int(d["ply"][i]), int(d["k"][i]), d["act"][i].copy(),
                                                                d["n"][i].copy(), int(d["choice"][i])))
            for g, ln, r in zip(d["fin_game"], d["fin_len"], d["fin_result"]):
                fin[int(g)] = (int(ln), int(r))
        eng.check_status()
        assert seen and sorted(fin) == list(range(6)), (sample_cap, fin_cap)
        for g in range(6):
            _assert_game_equals(sorted(plies[g], key=lambda p: p[0]), fin[g], want)
    # node pool too small for one search: flagged, never silent
    small = ChessTreeEngine(n_trees=2, sims_per_move=64, eval_mode="uniform", prior_mode="f64", max_plies=4, node_capacity=256)
    small.search()
    torch.cuda.synchronize()
    assert int(small.view("status")[0]) & native.AZ_FLAG_POOL_OVERFLOW
    with pytest.raises(NativeError, match="node pool exhausted"):
        small.check_status()


def test_stem_from_boards_equals_the_convolution_on_planes():
    """az_chess_stem (20 varying planes from the 64-byte board + per-cell constant) against cuDNN's stem convolution on
    the full 118-plane tensor, for positions of the self-play path (clocks, castling rights, promotions, en passant)."""
    from az_b200 import chess
    from az_b200.chess_selfplay import chess_net
    from az_b200.net import InferenceNet, randomise_bn

    torch.manual_seed(4)
    net = randomise_bn(chess_net()).eval()
    inf = InferenceNet(net, dtype=torch.bfloat16, device="cuda")
    acts = cr.all_possible_moves()
    index = {m: i for i, m in enumerate(acts)}
    samples = []
    for game in range(12):
        rng = _lcg_local(game)
        s = cr.start_state()
        for ply in range(120):
            moves = [m for m in cr.legal(s) if m in index]
            if cr.status(s) != 0 or not moves:
                break
            samples.append(s.copy())
            s = cr.push(s, moves[next(rng) % len(moves)], keep_same_player=True)
    pos = np.stack([cr.to_pos(x) for x in samples])
    assert len(pos) > 800 and any(x.halfmove > 5 for x in samples) and any(x.castling != 15 for x in samples)
    planes = chess.chess_encode(pos, dtype=torch.bfloat16)                       # [n, 8, 8, 118] bf16 on the GPU
    x = torch.nn.functional.pad(planes, (0, inf.in_pad)).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    ref = torch.cudnn_convolution_relu(x, inf.stem_w_pad, inf.stem_b, (1, 1), (1, 1), (1, 1), 1).permute(0, 2, 3, 1).float()
    dev_pos = torch.from_numpy(pos.view(np.int64)).cuda()
    for tc in (True, False):  # tcgen05 and mma.sync versions
        for n in (len(pos), 1, 7):  # odd counts: the last tcgen05 tile holds one position
            got = inf.chess_stem(dev_pos[:n], tc=tc).float()
            assert got.shape == (n, 8, 8, 128)
            err = (got - ref[:n]).abs().max().item()
            assert err <= 2.0 ** -6 * max(1.0, ref.abs().max().item()), (tc, n, err)   # one bf16 rounding, other summation order
    assert float(ref.max()) > 0.5 and float((ref > 0).float().mean()) > 0.2  # the probe is not all zeros after the ReLU
    # end to end: priors / values through both routes
    p1, v1 = inf.forward_from_stem(inf.chess_stem(torch.from_numpy(pos.view(np.int64)).cuda()))
    p0, v0 = inf(planes)
    assert (p1 - p0).abs().max().item() <= 2e-3 * max(p0.max().item(), 1e-3) + 1e-5 and (v1 - v0).abs().max().item() <= 2e-2


def test_runner_with_stem_from_boards_plays_the_same_games():
    """The two leaf routes (planes + cuDNN stem / boards + az_chess_stem) feed the same search: with argmax moves and a
    short horizon the games coincide unless a near-tie in the bf16 net flips a visit; at least most plies must agree."""
    from az_b200.chess_selfplay import ChessSelfPlayRunner, chess_net

    outs = []
    for flag in (False, True):
        torch.manual_seed(0)
        r = ChessSelfPlayRunner(n_trees=8, sims_per_move=16, net=chess_net(), games_target=8, max_plies=6, unroll=2,
                                move_mode="argmax", auto_restart=False, stem_from_boards=flag)
        assert r.stem_from_boards == flag
        states, policies, values, known = r.run_until_done(poll_every=16, max_advances=4000)
        assert states.shape == (48, 8, 8, 118) and known.all()
        outs.append(policies)
    agree = (np.abs(outs[0] - outs[1]).max(-1) < 0.13).mean()   # a flipped visit moves a target by 1/15
    assert agree >= 0.8, agree


def test_tail_planes_are_the_last_34_and_change_nothing():
    """az_chess_step with plane_first = 84 writes planes 84-117 (+ 6 zero channels); the net gives the same priors and
    values from them as from all 118 planes, because planes 0-83 are empty on the self-play path."""
    from az_b200 import chess
    from az_b200.chess_engine import ChessTreeEngine
    from az_b200.chess_selfplay import chess_net
    from az_b200.net import InferenceNet, randomise_bn

    torch.manual_seed(6)
    T = 64
    eng = ChessTreeEngine(n_trees=T, sims_per_move=20, eval_mode="external", move_mode="philox", seed=3, index_move_greedy=99)
    full = torch.full((T, 8, 8, 120), 7.0, dtype=torch.bfloat16, device="cuda")
    tail = torch.full((T, 8, 8, 40), 7.0, dtype=torch.bfloat16, device="cuda")
    valid = torch.zeros(T, dtype=torch.int32, device="cuda")
    pri = torch.rand(T, 1880, device="cuda")
    val = torch.zeros(T, device="cuda")
    inf = InferenceNet(randomise_bn(chess_net()).eval(), dtype=torch.bfloat16, device="cuda")
    checked = 0
    for adv in range(120):
        if adv % 2 == 0:
            eng.step(pri if adv else None, val if adv else None, full, valid)
        else:
            eng.step(pri, val, tail, valid, plane_first=84)
        if adv % 20 == 19:
            eng.move()
        ok = valid.bool()
        if not bool(ok.any()):
            continue
        pos = eng.view("leaf_pos")[ok].cpu().numpy().view(np.uint64)
        want = torch.from_numpy(chess.chess_encode(pos)).cuda()
        if adv % 2 == 0:
            assert torch.equal(full[ok][..., :118].float(), want) and not full[ok][..., 118:].any()
            assert not want[..., :84].any()  # the empty history entries
        else:
            assert torch.equal(tail[ok][..., :34].float(), want[..., 84:]) and not tail[ok][..., 34:].any()
            p_tail, v_tail = inf(tail[ok])
            p_full, v_full = inf(torch.nn.functional.pad(want, (0, 2)).to(torch.bfloat16))
            assert (p_tail - p_full).abs().max().item() <= 1e-3 * p_full.max().item() + 1e-6
            assert (v_tail - v_full).abs().max().item() <= 1e-2
            checked += int(ok.sum())
    assert checked > 1000


def test_drop_in_mcts_class_for_chess(monkeypatch):
    """custom_alphazero.mcts.mcts.MCTS with ConfigGeneral.game = "chess", driven like the reference drives it (search,
    play(return_details=True, deterministic=True), evaluator through the module-level infer_sample hook), against the
    C oracle's MCTS with the same fixed evaluator."""
    import importlib
    import sys

    from custom_alphazero import config

    monkeypatch.setattr(config.ConfigGeneral, "game", "chess")
    monkeypatch.setattr(config.ConfigMCTS, "index_move_greedy", 3)
    for name in ("custom_alphazero.mcts.mcts", "custom_alphazero.mcts.chess_mcts"):
        sys.modules.pop(name, None)
    mm = importlib.import_module("custom_alphazero.mcts.mcts")
    try:
        from custom_alphazero.chess.board import Board
        from custom_alphazero.chess.utils import get_all_possible_moves

        calls = []

        def uniform(state, concurrency):
            calls.append(state.shape)
            return np.full(1880, 1 / 1880), 0.0

        monkeypatch.setattr(mm, "infer_sample", uniform)
        sims, plies = 40, 6
        want = cr.mcts_game(sims=sims, evaluator="uniform", prior_mode="f64", max_plies=plies, greedy_idx=3)
        moves = get_all_possible_moves()
        search = mm.MCTS(board=Board(), all_possible_moves=moves, concurrency=False, plays_inferences={}, model=None)
        assert type(search).__name__ == "ChessMCTS" and search.current_root.edges == []
        for ply in range(plies):
            search.search(sims)
            k = int(want["k"][ply])
            edges = search.current_root.edges
            assert [e.visit_count for e in edges] == want["n"][ply][:k].tolist(), ply
            assert [moves.index(e.action) for e in edges] == want["act"][ply][:k].tolist()
            greedy = ply >= 3
            parent, child, policy, move = search.play(greedy, return_details=True, deterministic=True)
            assert moves.index(move) == int(want["choice"][ply]) & 0xFFFF
            assert parent.shape == child.shape == (8, 8, 118) and policy.shape == (1880,)
            n = want["n"][ply][:k].astype(np.float64)
            ref = np.zeros(1880)
            if greedy:
                ref[want["act"][ply][int(np.argmax(n))]] = 1.0
            else:
                ref[want["act"][ply][:k].astype(np.int64)] = n / n.sum()
            assert np.array_equal(policy, ref), ply
        assert calls and set(calls) == {(8, 8, 118)}
        assert search.board.turn is True and not search.board.is_game_over()
    finally:
        for name in ("custom_alphazero.mcts.mcts", "custom_alphazero.mcts.chess_mcts"):
            sys.modules.pop(name, None)


def _lcg_local(seed):
    s = (seed * 0x9E3779B97F4A7C15 + 1) % 2 ** 64
    while True:
        s = (s * 6364136223846793005 + 1442695040888963407) % 2 ** 64
        yield s >> 33


@pytest.mark.parametrize("name", ["chess_mcts_hash_f64_100", "chess_mcts_uniform_f32_64", "chess_mcts_hash_f32_200"])
def test_engine_reproduces_the_committed_fixtures(name):
    """tests/golden/chess_mcts_*.json (frozen oracle games, tests/golden/make_chess_golden.py) on the GPU."""
    import json
    import os

    from az_b200.chess_engine import ChessTreeEngine

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".json")) as fp:
        w = json.load(fp)
    eng = ChessTreeEngine(n_trees=3, sims_per_move=w["sims"], eval_mode=w["evaluator"], prior_mode=w["prior_mode"],
                          max_plies=w["max_plies"], index_move_greedy=w["greedy_idx"])
    per_game, fin = _play_games(eng, w["max_plies"] + 2)
    for g in range(3):
        assert fin[g] == (w["plies"], w["result"])
        for ply, k, act, n, choice in per_game[g]:
            assert k == w["k"][ply] and choice == w["choice"][ply]
            assert act[:k].tolist() == w["act"][ply] and n[:k].tolist() == w["n"][ply]
