"""CPU tests of the on-disk formats the reference's other processes consume: checkpoint
directory (weights + meta.json + sentinel), results/ path layout, samples.npz keys."""
import json
import os

import numpy as np
import pytest
import torch


def test_checkpoint_roundtrip_sentinel_and_hash(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from az_b200.net import PolicyValueNet, randomise_bn
    from custom_alphazero import paths
    from custom_alphazero.utils import (best_saved_model, best_saved_model_hash, best_saved_model_path, load_with_meta,
                                        model_hash, save_with_meta)

    run = "run-1"
    assert best_saved_model_path(run) is None and best_saved_model_hash(run) is None
    os.makedirs(paths.get_evaluation_path(run))  # Q8: an empty evaluation directory means "no model yet"
    assert best_saved_model_path(run) is None
    torch.manual_seed(3)
    net = randomise_bn(PolicyValueNet())
    for it in (0, 2, 10):
        save_with_meta(net, os.path.join(paths.get_evaluation_path(run), f"iteration_{it}"), steps=7 * it, learning_rate=1e-3)
    # a half-written checkpoint (no sentinel) is ignored even though its number is the largest
    os.makedirs(os.path.join(paths.get_evaluation_path(run), "iteration_11"))
    assert best_saved_model_path(run).endswith("iteration_10")  # numeric, not lexicographic, order
    files = set(os.listdir(best_saved_model_path(run)))
    assert {"model.pt", "meta.json", "MODEL_SAVED_SUCCESSFULLY"} <= files
    meta = json.load(open(os.path.join(best_saved_model_path(run), "meta.json")))
    assert meta == {"hash": model_hash(net), "learning_rate": 1e-3, "steps": 70}
    assert best_saved_model_hash(run) == model_hash(net)
    loaded = best_saved_model(run)
    assert model_hash(loaded) == model_hash(net)
    # corrupted weights are refused
    other = PolicyValueNet()
    torch.save(other.state_dict(), os.path.join(best_saved_model_path(run), "model.pt"))
    with pytest.raises(AssertionError):
        load_with_meta(PolicyValueNet(), best_saved_model_path(run))


def test_results_layout_matches_the_reference():
    from custom_alphazero import paths

    assert paths.get_self_play_samples_path("r", 3) == os.path.join("results", "connect_n", "r", "self_play", "iteration_3", "samples.npz")
    assert paths.get_training_path("r") == os.path.join("results", "connect_n", "r", "training")
    assert paths.get_evaluation_path("r") == os.path.join("results", "connect_n", "r", "evaluation")


def test_append_queue_payload_shape(monkeypatch):
    """PATCH /api/queue/append carries nested lists under states / policies / values (serving/factory.py:69-80)."""
    import custom_alphazero.serving.factory as f

    sent = {}

    class Resp:
        status_code = 200

    class FakeRequests:
        @staticmethod
        def patch(url, data, headers, timeout):
            sent.update(url=url, data=json.loads(data))
            return Resp()

    monkeypatch.setattr(f, "_requests", lambda: FakeRequests)
    ok = f.append_queue(np.zeros((2, 6, 7, 4), np.float32), np.full((2, 7), 1 / 7), np.asarray([1, -1]))
    assert ok and sent["url"].endswith("/api/queue/append")
    assert set(sent["data"]) == {"states", "policies", "values"} and sent["data"]["values"] == [1, -1]
    assert np.asarray(sent["data"]["states"]).shape == (2, 6, 7, 4)


def test_checkpoint_exported_by_the_reference_loads(tmp_path):
    """A reference-side export (np.savez of PolicyValueModel.get_weights() next to meta.json and the sentinel) is
    readable: the path from a reference checkpoint to PolicyValueNet (model.py:190-212)."""
    import numpy as np

    from az_b200.net import PolicyValueNet, randomise_bn
    from custom_alphazero.config import ConfigPath
    from custom_alphazero.utils import KERAS_WEIGHTS, load_with_meta

    torch.manual_seed(5)
    src = randomise_bn(PolicyValueNet()).eval()
    d = tmp_path / "iteration_3"
    d.mkdir()
    np.savez(d / KERAS_WEIGHTS, *src.to_keras_weights())  # what the maintainer runs on the reference side
    json.dump({"steps": 12, "learning_rate": 0.01, "hash": 1234567890123456789012345678901234567890}, open(d / ConfigPath.model_meta, "w"))
    with pytest.raises(AssertionError):  # no sentinel yet
        load_with_meta(PolicyValueNet(), str(d))
    open(d / ConfigPath.model_success, "wb").close()
    dst = PolicyValueNet().eval()
    meta = load_with_meta(dst, str(d))
    assert meta["steps"] == 12
    x = torch.zeros(2, 6, 7, 4)
    x[..., 0] = 1
    with torch.no_grad():
        assert torch.equal(src(x)[0], dst(x)[0]) and torch.equal(src(x)[1], dst(x)[1])
