"""GPU tests of the drop-in package (custom_alphazero.*): the reference's own call patterns, checked
against the golden vectors produced by the unmodified reference."""
import numpy as np
import pytest

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture()
def c4():
    from custom_alphazero.config import ConfigConnectN

    saved = (ConfigConnectN.board_width, ConfigConnectN.board_height, ConfigConnectN.n, ConfigConnectN.gravity)
    yield ConfigConnectN
    ConfigConnectN.board_width, ConfigConnectN.board_height, ConfigConnectN.n, ConfigConnectN.gravity = saved


def _configure(cfg, case):
    cfg.board_width, cfg.board_height, cfg.n, cfg.gravity = case["W"], case["H"], case["n"], case["gravity"]


def test_board_terminal_and_state_semantics(c4):
    """SURVEY 8c 'terminal / sign semantics' row, through the compat Board."""
    from custom_alphazero.connect_n.board import Board
    from custom_alphazero.connect_n.move import Move

    b = Board()
    for x in (0, 1, 0, 1, 0, 1):
        b.play(Move(True, x), keep_same_player=True)
    assert repr(b).split("\n")[3:] == ["XO.....", "XO.....", "XO....."]
    assert b.turn == 1 and b.fullmove_number == 6 and not b.is_game_over()
    fs = b.full_state
    assert fs.dtype == np.float32 and fs.shape == (6, 7, 4)
    assert fs.sum(axis=(0, 1)).tolist() == [36.0, 3.0, 3.0, 42.0]
    assert fs[5, 0].tolist() == [0, 1, 0, 1] and fs[5, 1].tolist() == [0, 0, 1, 1] and fs[0, 0].tolist() == [1, 0, 0, 1]
    child = b.play(Move(True, 0), on_copy=True, keep_same_player=True)
    assert child.is_game_over() and child.is_null is False and child.get_result(keep_same_player=True) == 1
    assert not b.is_game_over() and [str(m) for m in b.moves] == ["0", "1", "2", "3", "4", "5", "6"]
    assert child.play(Move(True, 3), on_copy=True) is child  # Q7
    mask = b.legal_moves_mask(Board.get_all_possible_moves())
    assert mask.dtype == bool and mask.all()
    with pytest.raises(AssertionError):
        full = Board(np.tile(np.array([[1], [-1], [1], [-1], [1], [-1]], dtype=np.int8), (1, 7)))
        full.push(Move(True, 2))


def test_board_without_keep_same_player_alternates_colours(c4):
    from custom_alphazero.connect_n.board import Board
    from custom_alphazero.connect_n.move import Move

    b = Board()
    b.play(Move(True, 3))
    b.play(Move(True, 3))
    assert b.array[5, 3] == 1 and b.array[4, 3] == -1 and b.turn == 1 and b.fullmove_number == 2
    assert b.full_state[:, :, 3].min() == 1.0
    b.play(Move(True, 0))
    assert b.turn == -1 and b.full_state[:, :, 3].max() == -1.0


@pytest.mark.parametrize("name", ["game_6x7_250_uniform", "game_6x7_250_hash_seed7", "game_5x5ng_n3_60_hash"])
def test_compat_mcts_reproduces_reference_game(name, c4):
    """Drives custom_alphazero.mcts.mcts.MCTS exactly like the golden generator drove the reference
    (infer_sample patched, model=None, play(greedy, return_details=True, deterministic=...))."""
    from oracle import evaluators

    case = load_golden(name)
    _configure(c4, case)
    import custom_alphazero.mcts.mcts as m
    from custom_alphazero.connect_n.board import Board

    all_moves = Board.get_all_possible_moves()
    A = len(all_moves)
    f = evaluators.make(case["evaluator"], A)
    saved = m.infer_sample
    m.infer_sample = lambda state, concurrency: f(state)
    try:
        if case.get("seed") is not None:
            np.random.seed(case["seed"])
        mcts = m.MCTS(Board(), all_moves, False, {}, model=None)
        t = 0
        while not mcts.board.is_game_over():
            mcts.search(case["sims"])
            want = case["plies"][t]
            edges = mcts.current_root.edges
            assert [e.visit_count for e in edges] == want["N"], f"ply {t}"
            assert [e.total_action_value for e in edges] == want["W"], f"ply {t}"
            assert [e.prior for e in edges] == want["P"], f"ply {t}"
            assert [all_moves.index(e.action) for e in edges] == want["actions"]
            greedy = mcts.board.fullmove_number >= 8
            parent_state, child_state, policy, move = mcts.play(greedy, return_details=True,
                                                                deterministic=case.get("seed") is None)
            assert all_moves.index(move) == want["move"] and str(move) == want["move_str"]
            assert policy.dtype == np.float64 and policy.tolist() == want["policy"]
            assert parent_state.dtype == np.float32 and parent_state.shape == (case["H"], case["W"], 4)
            assert child_state.shape == parent_state.shape
            t += 1
        assert t == case["n_plies"] and mcts.board.get_result(keep_same_player=True) == case["result"]
        assert repr(mcts.board) == case["final_repr"]
    finally:
        m.infer_sample = saved


def test_compat_mcts_model_hook_and_cache(c4):
    """model= hook (mcts.py:131-137): called on [1, H, W, 4], outputs with .numpy(); float32 priors are
    normalised in float32; the plays_inferences cache is keyed by the position text."""
    import custom_alphazero.mcts.mcts as m
    from custom_alphazero.connect_n.board import Board
    from oracle import c_oracle

    calls = []

    class Out:
        def __init__(self, a):
            self.a = a

        def numpy(self):
            return self.a

    table = np.asarray([0.3, 0.05, 0.2, 0.1, 0.15, 0.12, 0.08], dtype=np.float32)

    def model(x):
        assert x.shape == (1, 6, 7, 4) and x.dtype == np.float32
        calls.append(1)
        shift = int(x[0, :, :, 1].sum()) % 7
        return Out(np.roll(table, shift)[None]), Out(np.asarray([[0.25 * ((shift % 3) - 1)]], dtype=np.float32))

    cache = {}
    mcts = m.MCTS(Board(), Board.get_all_possible_moves(), False, cache, model=model)
    mcts.search(300)
    got = [e.visit_count for e in mcts.current_root.edges]

    def cb(state):
        shift = int(state[:, :, 1].sum()) % 7
        return np.roll(table, shift).astype(np.float64), float(np.float32(0.25 * ((shift % 3) - 1)))

    want = c_oracle.search_once(c_oracle.make_rules(7, 6, 4, True), [], 300, "callback", c_oracle.PRIOR_F32, cb)
    assert got == want["N"]
    assert [e.prior for e in mcts.current_root.edges] == want["P"]
    assert len(cache) == len(calls) and len(calls) <= 300
    assert all(set(k) <= set("XO.\n") for k in cache)


def test_tree_views_are_fetched_on_demand(c4):
    """UCTNode / UCTEdge views (mcts.py:22-105 of the reference are live Python objects): reading the root's edges copies
    the root's child block only, deeper nodes pull the live pool once, a view that was read keeps its values when the
    tree moves on, and a view nobody read refuses to describe a tree that no longer exists."""
    import custom_alphazero.mcts.mcts as m
    from custom_alphazero.connect_n.board import Board
    from oracle import evaluators

    f = evaluators.hash_evaluator(7)
    saved = m.infer_sample
    m.infer_sample = lambda state, concurrency=False: f(state)
    try:
        t = m.MCTS(Board(), Board.get_all_possible_moves(), False, {})
        unread = t.current_root
        t.search(60)
        with pytest.raises(RuntimeError):
            unread.edges
        root = t.current_root
        snap = root._export
        n60 = [e.visit_count for e in root.edges]
        assert sum(n60) == 59 and snap.full is None and snap.root_block is not None  # four small copies, no pool copy
        t.search(60)  # the view that was read is completed before the tree changes ...
        assert snap.full is not None and [e.visit_count for e in root.edges] == n60
        deep = root.edges[int(np.argmax(n60))].child
        assert sum(e.visit_count for e in deep.edges) == max(n60) - 1  # ... and still describes the 60-simulation tree
        n120 = [e.visit_count for e in t.current_root.edges]
        assert sum(n120) == 119 and all(b >= a for a, b in zip(n60, n120))
        grand = t.current_root.edges[int(np.argmax(n120))].child
        assert sum(e.visit_count for e in grand.edges) == max(n120) - 1 and t.current_root._export.full is not None
    finally:
        m.infer_sample = saved


def test_self_play_play_returns_reference_shaped_arrays(c4):
    from custom_alphazero import self_play
    from custom_alphazero.config import ConfigB200, ConfigSelfPlay

    saved = (ConfigB200.games_per_iteration, ConfigB200.concurrent_games, ConfigSelfPlay.mcts_iterations)
    ConfigB200.games_per_iteration, ConfigB200.concurrent_games, ConfigSelfPlay.mcts_iterations = 96, 64, 24
    try:
        states, policies, rewards, trees = self_play.play("test-run")
    finally:
        ConfigB200.games_per_iteration, ConfigB200.concurrent_games, ConfigSelfPlay.mcts_iterations = saved
    S = len(rewards)
    assert states.shape == (S, 6, 7, 4) and states.dtype == np.float32
    assert policies.shape == (S, 7) and policies.dtype == np.float64
    assert np.allclose(policies.sum(-1), 1.0) and set(np.unique(rewards)) <= {-1, 0, 1}
    assert (states[..., 3] == 1).all() and (states[..., :3].sum(-1) == 1).all()
    # first sample of every game is the empty board; there are 96 games
    assert int((states[..., 0].sum(axis=(1, 2)) == 42).sum()) == 96


def test_root_policy_targets_bf16_vs_fp32_within_tolerance(c4):
    """SURVEY 4 test (4): the reference MCTS API driven once with the fp32 CPU module and once with the bf16 GPU
    inference path (same weights), 400 simulations from random opening positions; root policy targets N / sum N.
    Tolerance: max |d pi| <= 2.5e-2, mean <= 5e-3.  Measured on B200: max 5e-3 (two visits of 399), mean 1.9e-3,
    5 of 12 positions identical (tools/explore_rootpi.py).  One flipped selection moves a visit = 2.5e-3, so
    north_star's illustrative 1e-3 is below the resolution of a 400-simulation search with near-tied PUCT scores."""
    import custom_alphazero.mcts.mcts as m
    from az_b200 import net as N
    from custom_alphazero.connect_n.board import Board

    torch.manual_seed(0)
    ref = N.randomise_bn(N.PolicyValueNet()).eval()
    inf = N.InferenceNet(ref, device="cuda")

    def cpu_model(x):
        with torch.no_grad():
            return ref(torch.from_numpy(x))

    def gpu_model(x):
        p, v = inf(torch.from_numpy(x).cuda().to(torch.bfloat16))
        return p.cpu(), v.cpu()

    all_moves = Board.get_all_possible_moves()
    rng = np.random.RandomState(1)
    diffs = []
    for _ in range(6):
        b = Board()
        for _ in range(rng.randint(0, 8)):
            mv = b.moves
            if b.is_game_over():
                break
            b.play(mv[rng.randint(len(mv))], keep_same_player=True)
        if b.is_game_over():
            continue
        pis = []
        for model in (cpu_model, gpu_model):
            t = m.MCTS(b, all_moves, False, {}, model=model)
            t.search(400)
            n = np.asarray([e.visit_count for e in t.current_root.edges], dtype=np.float64)
            assert n.sum() == 399
            pis.append(n / n.sum())
        diffs.append(np.abs(pis[0] - pis[1]).max())
    assert len(diffs) >= 4 and max(diffs) <= 2.5e-2 and np.mean(diffs) <= 5e-3, diffs


def test_self_play_entry_point_writes_reference_format(tmp_path, monkeypatch, c4):
    """`python -m custom_alphazero.self_play` (here: its main(), one iteration) without a serving process:
    stand-alone run id, results/connect_n/{run}/self_play/iteration_0/samples.npz with the reference's keys,
    drawn games excluded."""
    import os

    from custom_alphazero import self_play
    from custom_alphazero.config import ConfigB200, ConfigSelfPlay

    monkeypatch.chdir(tmp_path)
    saved = (ConfigB200.games_per_iteration, ConfigB200.concurrent_games, ConfigSelfPlay.mcts_iterations)
    ConfigB200.games_per_iteration, ConfigB200.concurrent_games, ConfigSelfPlay.mcts_iterations = 40, 32, 16
    try:
        self_play.main(max_iterations=1)
    finally:
        ConfigB200.games_per_iteration, ConfigB200.concurrent_games, ConfigSelfPlay.mcts_iterations = saved
    runs = os.listdir(os.path.join("results", "connect_n"))
    assert len(runs) == 1 and runs[0].startswith("standalone_")
    path = os.path.join("results", "connect_n", runs[0], "self_play", "iteration_0", "samples.npz")
    data = np.load(path)
    assert set(data.files) == {"states", "policies", "values"}
    S = len(data["values"])
    assert S > 100 and data["states"].shape == (S, 6, 7, 4) and data["states"].dtype == np.float32
    assert data["policies"].shape == (S, 7) and data["policies"].dtype == np.float64
    assert set(np.unique(data["values"])) <= {-1, 1}  # exclude_null_games: no zero rewards left


def test_integration_md_binding_snippet_runs(c4):
    """The ctypes stub printed in INTEGRATION.md section 3 is executed as written (only the library path is made
    absolute): it must create an engine and drive az_step / az_play with a stand-in evaluator."""
    import ctypes
    import os
    import re

    from tests.helpers import ROOT

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "class AzConfig" in b)
    lib_path = os.path.join(ROOT, "custom-alphazero_b200", "libaz_b200.so")
    stub = stub.replace('ctypes.CDLL("libaz_b200.so")', f'ctypes.CDLL("{lib_path}")')
    ns = {}
    exec(stub, ns)
    assert ctypes.sizeof(ns["AzConfig"]) > 0
    T, sims = 16, 12
    h, slab, lay, cfg = ns["make_engine"](None, T, sims)
    lib = ns["_lib"]
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    A = lay.n_actions
    priors = torch.full((T, A), 1.0 / A, dtype=torch.float32, device="cuda")
    values = torch.zeros(T, dtype=torch.float32, device="cuda")
    states = torch.zeros((T, 6, 7, 4), dtype=torch.bfloat16, device="cuda")
    valid = torch.zeros(T, dtype=torch.int32, device="cuda")
    for _ in range(400):
        assert lib.az_step(h, ptr(priors), ptr(values), 0, ptr(states), 2, ptr(valid), stream) == 0
        assert lib.az_play(h, -1, -1, stream) == 0
    torch.cuda.synchronize()
    counters = slab[lay.counters: lay.counters + T * 8 * 8].view(torch.int64).view(T, 8).sum(0).tolist()
    status = slab[lay.status: lay.status + 4 * T].view(torch.int32)
    assert counters[0] >= 400 * T and counters[2] >= (400 // sims - 1) * T and counters[3] > 0  # sims, moves, games
    assert int((status & ~0xFF).max()) == 0  # no error flags
    st = states.float()
    assert bool((st[..., 3] == 1).all()) and bool((st[..., :3].sum(-1) == 1).all())


_MULTI_RANK_ENTRY = r'''
import os, sys
sys.path.insert(0, os.path.join(sys.argv[1], "custom-alphazero_b200"))
os.chdir(sys.argv[2])
os.environ["AZ_DIST_BACKEND"] = "gloo"   # two ranks share the one GPU of the test box: NCCL refuses that, gloo does not
import torch
torch.manual_seed(0)                      # every rank builds the same random-init net (no checkpoint in a stand-alone run)
from custom_alphazero import self_play
from custom_alphazero.config import ConfigB200, ConfigSelfPlay
ConfigB200.games_per_iteration, ConfigB200.concurrent_games, ConfigSelfPlay.mcts_iterations = 40, 32, 16
self_play.main(max_iterations=1)
print("rank", os.environ.get("RANK", "0"), "ok")
'''


def test_self_play_entry_point_on_two_ranks_collects_the_single_rank_games(tmp_path, c4):
    """torchrun --nproc-per-node 2 -m custom_alphazero.self_play (SURVEY 4 test 5 for the product entry point): the 40
    games of an iteration sharded over two ranks and gathered to rank 0 must be, sample for sample, the games one rank
    plays alone - a game depends on its id, the seed and the weights, not on the rank or the batch it shares."""
    import os
    import subprocess
    import sys

    from tests.helpers import ROOT

    script = tmp_path / "entry.py"
    script.write_text(_MULTI_RANK_ENTRY)
    data = []
    for world, port in ((1, 29621), (2, 29622)):
        work = tmp_path / f"w{world}"
        work.mkdir()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), str(script), ROOT, str(work)]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
        assert out.stdout.count("ok") == world
        runs = os.listdir(work / "results" / "connect_n")
        assert len(runs) == 1
        data.append(np.load(work / "results" / "connect_n" / runs[0] / "self_play" / "iteration_0" / "samples.npz"))
    one, two = data
    assert len(one["values"]) > 100
    for k in ("states", "policies", "values"):
        assert np.array_equal(one[k], two[k]), k
