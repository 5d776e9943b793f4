"""Multi-GPU plumbing: one process per GPU, games sharded by id, no collective on the simulation path.

The reference spreads games over os.cpu_count()-1 worker processes (self_play.py:98-110), "broadcasts"
weights by re-reading a checkpoint directory (self_play.py:142-150, utils.py:64-78) and ships samples as
JSON over HTTP (serving/factory.py:69-80).  Here: contiguous game-id ranges per rank, torch.distributed
broadcast of the flat weight vector, all-gather of the compact game records.  Works with the nccl
backend (CUDA tensors) and the gloo backend (CPU tensors, used by the CPU tests).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_games(total_games, rank=None, world_size=None):
    """Contiguous game-id range of a rank: (first id, count).  Results of a game depend only on its id
    (seed) and the weights, so the multiset of games is the same for any number of ranks."""
    if rank is None or world_size is None:
        rank, world_size = world()
    q, r = divmod(int(total_games), world_size)
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def broadcast_weights(flat, src=0):
    """Broadcasts a flat float tensor of weights from the trainer rank (in place)."""
    _, ws = world()
    if ws > 1:
        dist.broadcast(flat, src=src)
    return flat


def all_gather_records(records):
    """records: dict of tensors whose first dimension is the number of finished games on this rank
    (game_id, len, result, visits, action, board).  Returns the same dict holding every rank's games
    (each rank gets all of them; the trainer rank is the one that uses them)."""
    rank, ws = world()
    if ws == 1:
        return records
    first = next(iter(records.values()))
    n = torch.tensor([first.shape[0]], dtype=torch.int64, device=first.device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n)
    counts = [int(c) for c in counts]
    m = max(counts)
    out = {}
    for key, t in records.items():
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(ws)]
        dist.all_gather(parts, pad)
        out[key] = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    return out


CHESS_SAMPLE_KEYS = ("game", "ply", "pos", "k", "act", "n", "choice")
CHESS_FIN_KEYS = ("fin_game", "fin_len", "fin_result")


def all_gather_chess_rings(drained, device=None):
    """The chess runner's drained rings (numpy arrays: one entry per finished ply under CHESS_SAMPLE_KEYS, one per finished
    game under CHESS_FIN_KEYS) from every rank, concatenated in rank order - the replay-buffer gather for chess.  Game ids
    are global (game_id_base per rank), so the host join by game id (chess_engine.sample_values) works on the result."""
    import numpy as np

    rank, ws = world()
    if ws == 1:
        return drained
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    out = {}
    for keys in (CHESS_SAMPLE_KEYS, CHESS_FIN_KEYS):
        group = {}
        for k in keys:
            a = np.ascontiguousarray(drained[k])
            if a.dtype == np.uint64:
                a = a.view(np.int64)      # same bits: the collectives have no unsigned 64-bit type
            elif a.dtype == np.uint16:
                a = a.astype(np.int32)    # ... and gloo no 16-bit one: widened for the wire
            group[k] = torch.from_numpy(a).to(device)
        gathered = all_gather_records(group)
        for k in keys:
            g = gathered[k].cpu().numpy()
            out[k] = g.view(np.uint64) if drained[k].dtype == np.uint64 else g.astype(drained[k].dtype)
    return out
