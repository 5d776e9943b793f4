"""File-system glue the self-play entry point needs (reference utils.py:24-133): which weights to play
with.  Checkpoints of the B200 build are torch state dicts (`model.pt`) next to the reference's
meta.json / MODEL_SAVED_SUCCESSFULLY sentinel.  TensorFlow's checkpoint files cannot be parsed here (no TensorFlow), but the
reference's weights can: `np.savez(os.path.join(path, "model_keras.npz"), *model.get_weights())` on the reference side
gives a file load_with_meta() reads through PolicyValueNet.load_keras_weights (the get_weights() order and the
HWIO / [in, out] layouts are restated in az_b200/net.py and checked against oracle/net_ref.py)."""
import hashlib
import json
import os
from typing import Optional

import numpy as np
import torch

from az_b200.net import PolicyValueNet
from custom_alphazero import paths
from custom_alphazero.config import ConfigConnectN, ConfigModel, ConfigPath
from custom_alphazero.connect_n.board import Board


KERAS_WEIGHTS = "model_keras.npz"  # np.savez(path, *PolicyValueModel.get_weights()): arrays arr_0, arr_1, ...


def model_hash(net: PolicyValueNet) -> str:
    """Hash of the weights (the reference sums md5 digests of the printed weights, model.py:168-173; here md5 over
    the raw parameter and buffer bytes in state-dict order)."""
    h = hashlib.md5()
    for name, t in net.state_dict().items():
        h.update(name.encode())
        h.update(t.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def save_with_meta(net: PolicyValueNet, path: str, steps: int = 0, learning_rate: float = ConfigModel.maximum_learning_rate):
    """Checkpoint layout of the reference (model.py:203-212): weights under `model*`, meta.json with steps,
    learning_rate and hash, and the MODEL_SAVED_SUCCESSFULLY sentinel written last so a half-written checkpoint
    is never picked up (utils.py:53,113-118)."""
    os.makedirs(path, exist_ok=True)
    torch.save(net.state_dict(), os.path.join(path, ConfigPath.model_prefix + ".pt"))
    meta = {"steps": int(steps), "learning_rate": float(learning_rate), "hash": model_hash(net)}
    with open(os.path.join(path, ConfigPath.model_meta), "w") as fp:
        json.dump(meta, fp, sort_keys=True, indent=4)
    open(os.path.join(path, ConfigPath.model_success), "wb").close()


def load_with_meta(net: PolicyValueNet, path: str) -> dict:
    """model.py:190-201: refuses a checkpoint without the sentinel or whose weights do not match the stored hash."""
    assert os.path.exists(os.path.join(path, ConfigPath.model_success)), f"No verification file of the model found at {path}!"
    with open(os.path.join(path, ConfigPath.model_meta)) as fp:
        meta = json.load(fp)
    pt = os.path.join(path, ConfigPath.model_prefix + ".pt")
    if not os.path.exists(pt) and os.path.exists(os.path.join(path, KERAS_WEIGHTS)):
        # a checkpoint exported by the reference: its hash is Keras' (md5 of the printed arrays), not comparable with ours
        with np.load(os.path.join(path, KERAS_WEIGHTS)) as z:
            net.load_keras_weights([z[k] for k in sorted(z.files, key=lambda k: int(k.split("_")[1]))])
        return meta
    net.load_state_dict(torch.load(pt, map_location="cpu"))
    assert model_hash(net) == meta.get("hash"), f"Unexpected weights hash recovered during model loading at {path}!"
    return meta


def init_model(path: Optional[str] = None) -> PolicyValueNet:
    from custom_alphazero.config import ConfigGeneral

    if ConfigGeneral.game == "chess":  # 8x8x118 planes in, one output per entry of chess.utils.get_all_possible_moves()
        from az_b200.chess import N_ACTIONS, PLANES

        net = PolicyValueNet(8, 8, N_ACTIONS, ConfigModel.filters, ConfigModel.depth, in_planes=PLANES)
    else:
        A = len(Board.get_all_possible_moves())
        net = PolicyValueNet(ConfigConnectN.board_height, ConfigConnectN.board_width, A, ConfigModel.filters, ConfigModel.depth)
    if path is not None:
        load_with_meta(net, path)
    return net.eval()


def _finished_iterations(run_id: str):
    root = paths.get_evaluation_path(run_id)
    if not os.path.isdir(root):
        return []
    done = [d for d in os.listdir(root) if os.path.exists(os.path.join(root, d, ConfigPath.model_success))]
    return sorted(done, key=lambda d: int(d.split("_")[-1]) if d.split("_")[-1].isdigit() else -1)


def best_saved_model_path(run_id: str) -> Optional[str]:
    """Newest finished evaluation iteration, or None (Q8: an empty evaluation directory is 'no model yet')."""
    done = _finished_iterations(run_id)
    return os.path.join(paths.get_evaluation_path(run_id), done[-1]) if done else None


def best_saved_model(run_id: str) -> PolicyValueNet:
    path = best_saved_model_path(run_id)
    if path is None:
        torch.manual_seed(0)
    return init_model(path)


def best_saved_model_hash(run_id: str):
    path = best_saved_model_path(run_id)
    if path is None:
        return None
    with open(os.path.join(path, ConfigPath.model_meta)) as fp:
        return json.load(fp).get("hash")


def reset_plays_inferences_dict() -> dict:
    return {}


class HostModel:
    """Adapter with the reference's model call convention (mcts.py:131-137): called on
    np.ndarray [1, H, W, 4], returns two tensors with .numpy() - evaluated by the bf16 GPU net."""

    def __init__(self, net: PolicyValueNet):
        from az_b200.net import InferenceNet

        self.inference = InferenceNet(net, device="cuda")

    def __call__(self, x):
        p, v = self.inference(torch.as_tensor(x, device="cuda"))
        return p.cpu(), v.cpu()
