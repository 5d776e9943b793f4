"""Times the three chess stems at 4096 positions: cuDNN on the 120-plane tensor, az_chess_stem (mma.sync) and
az_chess_stem_tc (tcgen05) from the 64-byte boards."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import numpy as np, torch
from az_b200 import chess
from az_b200.chess_selfplay import chess_net
from az_b200.net import InferenceNet
torch.backends.cudnn.benchmark = True
inf = InferenceNet(chess_net(), dtype=torch.bfloat16, device="cuda")
n = 4096
pos = torch.from_numpy(np.tile(chess.position_from_fen()[None].view(np.int64), (n, 1))).cuda()
planes = torch.nn.functional.pad(chess.chess_encode(pos, dtype=torch.bfloat16), (0, 2))
x = planes.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
def t(fn, rep=30):
    for _ in range(5): fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rep): fn()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / rep * 1e3
print("cudnn stem on planes  : %.1f us" % t(lambda: torch.cudnn_convolution_relu(x, inf.stem_w_pad, inf.stem_b, (1, 1), (1, 1), (1, 1), 1)))
print("az_chess_stem (mma)   : %.1f us" % t(lambda: inf.chess_stem(pos, tc=False)))
print("az_chess_stem_tc      : %.1f us" % t(lambda: inf.chess_stem(pos, tc=True)))
