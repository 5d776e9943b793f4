"""CPU tests of the host-side logic: sharding, collectives over gloo (world size 2), the fp32 net
restatement, the drop-in Move / normalize_probabilities / config."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.helpers import ROOT


def test_shard_games_partitions_exactly():
    from az_b200.dist import shard_games

    for total in (0, 1, 7, 4096, 32768, 32771):
        for ws in (1, 2, 3, 4, 8):
            parts = [shard_games(total, r, ws) for r in range(ws)]
            assert sum(c for _, c in parts) == total
            nxt = 0
            for base, count in parts:
                assert base == nxt
                nxt += count
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(sys.argv[1], "custom-alphazero_b200"))
import torch, torch.distributed as dist
from az_b200 import dist as azdist
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
flat = torch.arange(10, dtype=torch.float32) * (1.0 if rank == 0 else -1.0)
azdist.broadcast_weights(flat, src=0)
assert torch.equal(flat, torch.arange(10, dtype=torch.float32))
base, count = azdist.shard_games(11, rank, ws)
rec = {"game_id": torch.arange(base, base + count), "len": torch.full((count,), 3 + rank, dtype=torch.int32),
       "visits": torch.full((count, 4, 7), rank, dtype=torch.int32)}
allrec = azdist.all_gather_records(rec)
assert allrec["game_id"].tolist() == list(range(11)), allrec["game_id"]
assert allrec["visits"].shape == (11, 4, 7)
assert allrec["len"].tolist() == [3] * 6 + [4] * 5
assert allrec["visits"][:6].eq(0).all() and allrec["visits"][6:].eq(1).all()
# the trainer-only gather: rank 0 gets everything, the others nothing
one = azdist.gather_records(rec, dst=0)
if rank == 0:
    assert one["game_id"].tolist() == list(range(11)) and one["visits"].shape == (11, 4, 7)
    assert one["len"].dtype == torch.int32 and one["len"].tolist() == [3] * 6 + [4] * 5
    assert one["visits"][:6].eq(0).all() and one["visits"][6:].eq(1).all()
else:
    assert one is None
empty = {"game_id": torch.arange(0, 3 if rank == 1 else 0), "board": torch.ones((3 if rank == 1 else 0, 2, 5), dtype=torch.int8)}
one = azdist.gather_records(empty, dst=0)
if rank == 0:
    assert one["game_id"].tolist() == [0, 1, 2] and one["board"].shape == (3, 2, 5) and one["board"].eq(1).all()
# weight broadcast through the staging buffer: nothing lands in the live weights before take()
wb = azdist.WeightBroadcaster(10, "cpu")
live = torch.zeros(10)
assert not wb.take(live)
wb.start(torch.arange(10, dtype=torch.float32) + 5 if rank == 0 else None)
assert live.eq(0).all()
assert wb.take(live) and torch.equal(live, torch.arange(10, dtype=torch.float32) + 5)
assert not wb.take(live)
# chess: the two rings (finished plies, finished games) gathered separately, unsigned dtypes preserved
import numpy as np
n, f = 3 + rank, 1 + rank
d = {"game": np.full(n, 100 * rank, dtype=np.int64), "ply": np.arange(n, dtype=np.int32),
     "pos": np.full((n, 8), 2**63 + rank, dtype=np.uint64), "k": np.full(n, 20, dtype=np.int32),
     "act": np.full((n, 224), 0xffff - rank, dtype=np.uint16), "n": np.full((n, 224), rank, dtype=np.int32),
     "choice": np.full(n, 7, dtype=np.int32), "fin_game": np.full(f, 100 * rank, dtype=np.int64),
     "fin_len": np.full(f, n, dtype=np.int32), "fin_result": np.full(f, rank, dtype=np.int32)}
g = azdist.all_gather_chess_rings(d)
assert g["game"].tolist() == [0] * 3 + [100] * 4 and g["fin_game"].tolist() == [0, 100, 100]
assert g["pos"].dtype == np.uint64 and g["pos"][0, 0] == 2**63 and g["pos"][3, 0] == 2**63 + 1
assert g["act"].dtype == np.uint16 and g["act"][0, 0] == 0xffff and g["act"][6, 223] == 0xfffe
assert g["ply"].tolist() == [0, 1, 2, 0, 1, 2, 3] and g["fin_len"].tolist() == [3, 4, 4]
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_collectives_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_fp32_net_matches_the_reference_architecture_numbers():
    from az_b200 import net

    m = net.PolicyValueNet()
    assert m.n_parameters() == 1267037
    assert net.flops_per_eval(6, 7, 7) == 105037976
    assert net.flops_per_eval(9, 9, 9) == 202573412 and net.flops_per_eval(9, 9, 81) == 202596740
    x = torch.zeros(3, 6, 7, 4)
    x[..., 0] = 1
    x[..., 3] = 1
    p, v = m.eval()(x)
    assert p.shape == (3, 7) and v.shape == (3, 1) and torch.allclose(p.sum(-1), torch.ones(3))


def test_folded_inference_net_equals_module_on_cpu():
    from az_b200 import net

    torch.manual_seed(0)
    m = net.randomise_bn(net.PolicyValueNet(9, 9, 81)).eval()
    inf = net.InferenceNet(m, dtype=torch.float32, device="cpu")
    g = torch.Generator().manual_seed(1)
    code = torch.randint(0, 3, (8, 9, 9), generator=g)
    x = torch.zeros(8, 9, 9, 4).scatter_(3, code[..., None], 1.0)
    x[..., 3] = 1
    with torch.no_grad():
        p, v = m(x)
    p2, v2 = inf(x)
    assert (p - p2).abs().max() < 1e-5 and (v.reshape(-1) - v2).abs().max() < 1e-5
    flat = inf.flat_weights()
    inf.load_from(m)
    assert torch.equal(flat, inf.flat_weights())


def test_move_semantics():
    from custom_alphazero.connect_n.move import Move

    assert str(Move(True, 3)) == "3" and repr(Move(False, 2, 5)) == "(2, 5)"
    assert Move(True, 1) == Move(True, 1) and Move(False, 1, 2) != Move(False, 2, 1)
    assert sorted([Move(False, 2, 0), Move(False, 1, 5), Move(False, 1, 3)]) == [Move(False, 1, 3), Move(False, 1, 5), Move(False, 2, 0)]
    assert hash(Move(False, 4, 2)) == hash((4, 2)) and Move(True, 0) < Move(True, 6)
    with pytest.raises(AssertionError):
        Move(True, 1, 2)
    with pytest.raises(AssertionError):
        Move(False, 1)


def test_normalize_probabilities_semantics():
    from custom_alphazero.mcts.utils import normalize_probabilities

    np.testing.assert_array_equal(normalize_probabilities(np.zeros(4)), np.full(4, 0.25))
    p = np.full(7, 1 / 7)
    assert normalize_probabilities(p)[0] == 0.14285714285714288  # SURVEY 8c golden prior
    p32 = np.asarray([0.1, 0.2, 0.3], dtype=np.float32)
    assert normalize_probabilities(p32).dtype == np.float32
    with pytest.raises(AssertionError):
        normalize_probabilities(np.zeros(0))


def test_config_names_and_defaults():
    from custom_alphazero import config as c

    assert (c.ConfigConnectN.board_width, c.ConfigConnectN.board_height, c.ConfigConnectN.n, c.ConfigConnectN.gravity) == (7, 6, 4, True)
    assert c.ConfigSelfPlay.mcts_iterations == 250 and c.ConfigMCTS.exploration_constant == 1.5
    assert c.ConfigMCTS.index_move_greedy == 8 and c.ConfigSelfPlay.exclude_null_games is True
    assert c.ConfigConnectN.pieces == {-1: "O", 0: ".", 1: "X"} and c.ConfigServing.serving_port == 5555


def test_pow_half_table_is_cpython_pow():
    from az_b200.engine import pow_half_table

    t = pow_half_table(3000)
    assert t[2921] == 2921**0.5 and t[2921] != float(np.sqrt(np.float64(2921)))
    assert t[0] == 0.0 and t[1] == 1.0 and t[4] == 2.0


_ENTRY_WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(sys.argv[1], "custom-alphazero_b200"))
os.chdir(sys.argv[2])
os.environ["AZ_DIST_BACKEND"] = "gloo"
import numpy as np, torch
from az_b200.engine import Rules
from custom_alphazero import self_play
from custom_alphazero.config import ConfigB200, ConfigSelfPlay
import az_b200.selfplay as sp

ConfigB200.games_per_iteration = 11
seen = {}

class FakeRunner:  # stands in for the CUDA engine: every game of the shard "finishes" with 2 + id % 3 plies
    def __init__(self, games, base):
        self.rules, self.games, self.base = Rules(7, 6, 4, True), games, base
        seen["shard"] = (base, games)
    def run_until_done(self):
        pass
    def finished_device(self):
        ids = torch.arange(self.base, self.base + self.games)
        return {"game_id": ids, "len": (2 + ids % 3).to(torch.int32), "result": (ids % 2).to(torch.int32) * 2 - 1}
    def fin_clear(self):
        pass

def fake_decode(rules, fin, exclude_null_games=False, with_distance=False):
    order = torch.argsort(fin["game_id"])
    lens, res, ids = fin["len"][order].numpy(), fin["result"][order].numpy(), fin["game_id"][order].numpy()
    S = int(lens.sum())
    values = np.concatenate([np.full(l, r) for l, r in zip(lens, res)]).astype(np.int64)
    states = np.concatenate([np.full((l, 6, 7, 4), g, np.float32) for l, g in zip(lens, ids)])
    out = (states, np.full((S, 7), 1 / 7), values)
    return out + (np.concatenate([np.arange(l)[::-1] for l in lens]),) if with_distance else out

self_play._runner = lambda net, games, game_id_base: FakeRunner(games, game_id_base)
self_play.best_saved_model = lambda run_id: None
sp.decode_samples = fake_decode
self_play.main(max_iterations=2)
import torch.distributed as dist
rank = dist.get_rank()
base, count = seen["shard"]
assert (base, count) == ((11, 6) if rank == 0 else (17, 5)), (rank, base, count)   # iteration 1: ids 11..21 sharded 6 + 5
root = os.path.join("results", "connect_n")
runs = os.listdir(root)
assert len(runs) == 1
if rank == 0:
    for it in (0, 1):
        d = np.load(os.path.join(root, runs[0], "self_play", "iteration_%d" % it, "samples.npz"))
        ids = np.unique(d["states"][:, 0, 0, 0]).astype(int).tolist()
        assert ids == list(range(11 * it, 11 * it + 11)), ids        # every rank's games reached rank 0, none twice
        assert len(d["values"]) == sum(2 + g % 3 for g in ids)
print("rank", rank, "ok")
'''


def test_self_play_entry_point_world_size_2_gloo(tmp_path):
    """torchrun --nproc-per-node 2 -m custom_alphazero.self_play, host logic only (the CUDA engine and the decode kernel
    are stubbed): games sharded by id, records gathered to rank 0, which alone writes samples.npz; the run id is
    broadcast so that both ranks work in the same run directory.  Replaces the reference's joblib fan-out
    (self_play.py:98-110) and checkpoint polling (:142-150)."""
    script = tmp_path / "entry.py"
    script.write_text(_ENTRY_WORKER)
    work = tmp_path / "work"
    work.mkdir()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", str(script), ROOT, str(work)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
