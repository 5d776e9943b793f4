"""Shared helpers for the parity tests."""
import hashlib
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
MASK64 = (1 << 64) - 1


def golden_names(prefix):
    return sorted(f[:-5] for f in os.listdir(GOLDEN_DIR) if f.startswith(prefix) and f.endswith(".json"))


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as fp:
        return json.load(fp)


def trace_sha16(plies):
    """SURVEY 8c trace string: "{move}:{comma-joined root visit counts};" per ply."""
    s = "".join("{}:{};".format(p["move_str"], ",".join(map(str, p["N"]))) for p in plies)
    return hashlib.sha256(s.encode()).hexdigest()[:16]


def lcg_start(game):
    return (game * 0x9E3779B97F4A7C15 + 1) & MASK64


def lcg_next(s):
    return (s * 6364136223846793005 + 1442695040888963407) & MASK64
