"""Steady-state driver for ncu on the chess path: warm-up under graphs, then two eager advances inside an NVTX range
(ncu --nvtx --nvtx-include "steady_advances/"), then one perft launch (az_chess_perft)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import torch
from az_b200 import chess
from az_b200.chess_selfplay import ChessSelfPlayRunner
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
T = int(os.environ.get("TREES", 4096))
r = ChessSelfPlayRunner(n_trees=T, sims_per_move=800, games_target=1 << 40, unroll=8, sample_capacity=64 * T)
r.run(int(os.environ.get("WARM", 2400))); torch.cuda.synchronize()
r.engine.rings_clear()
for _ in range(4):
    r._advance()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("steady_advances")
for _ in range(2):
    r._advance()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
front = torch.from_numpy(chess.position_from_fen()[None].view("int64")).cuda()
for _ in range(4):
    mask, _, _ = chess.chess_legal(front)
    idx, act = torch.nonzero(mask, as_tuple=True)
    front, _ = chess.chess_play(front[idx], act.to(torch.int32))
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("perft")
n = chess.chess_perft(front, 3)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done", r.totals(), int(n.sum()))
