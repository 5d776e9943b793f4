"""HTTP client of the reference's serving process: only the calls the self-play entry point makes
(serving/factory.py:21-80 in the reference).  The control plane itself (FastAPI server, trainer) is out
of scope; these keep the wire format so a GPU self-play process can feed an unchanged trainer."""
import json
import uuid
from typing import Optional, Tuple

import numpy as np

from custom_alphazero.config import ConfigPath, ConfigServing


def _requests():
    import requests

    return requests


def infer_sample(state: np.ndarray, concurrency: bool) -> Tuple[np.ndarray, float]:
    """POST /api/inference: float64 priors + Python-float value; all-zero priors on a decode error."""
    requests = _requests()
    payload = {"uid": str(uuid.uuid4()), "state": state.tolist(), "concurrency": concurrency}
    url = ConfigServing.serving_address + "/api/inference"
    headers = {"content-type": "application/octet-stream"}
    try:
        resp = requests.post(url=url, data=json.dumps(payload), headers=headers, timeout=ConfigServing.inference_timeout)
    except requests.Timeout:
        payload["concurrency"] = False
        resp = requests.post(url=url, data=json.dumps(payload), headers=headers)
    try:
        content = json.loads(resp.content)
    except json.decoder.JSONDecodeError:
        from custom_alphazero.connect_n.board import Board

        content = {"probabilities": [0.0] * len(Board.get_all_possible_moves()), "value": 0.0}
    return np.asarray(content["probabilities"]), content["value"]


def get_run_id() -> Optional[str]:
    """GET /api/run-id; None when the server cannot be reached (the entry point then runs stand-alone)."""
    requests = _requests()
    try:
        resp = requests.get(url=ConfigServing.serving_address + ConfigPath.run_id_path, timeout=2)
        return json.loads(resp.content).get("run_id")
    except Exception:
        return None


def append_queue(states: np.ndarray, policies: np.ndarray, values: np.ndarray) -> bool:
    """PATCH /api/queue/append with nested lists, like the reference; False when there is no server."""
    requests = _requests()
    data = {"states": states.tolist(), "policies": policies.tolist(), "values": values.tolist()}
    try:
        resp = requests.patch(url=ConfigServing.serving_address + ConfigPath.append_queue_path, data=json.dumps(data),
                              headers={"content-type": "application/json"}, timeout=30)
        return resp.status_code == 200
    except Exception:
        return False
