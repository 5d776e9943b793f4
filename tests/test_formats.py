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
