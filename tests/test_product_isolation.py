"""The product must not route through the oracle or the reference (DESIGN.md 'Oracle')."""
import os
import re

from tests.helpers import ROOT

PRODUCT = os.path.join(ROOT, "custom-alphazero_b200")


def _sources(top, exts=(".py", ".cu", ".cuh", ".h", ".sh")):
    for d, _, files in os.walk(top):
        if "__pycache__" in d:
            continue
        for f in files:
            if f.endswith(exts):
                yield os.path.join(d, f)


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for path in _sources(PRODUCT):
        assert not pat.search(open(path).read()), f"{path} imports oracle/"


def test_nothing_that_travels_reads_the_reference_at_run_time():
    allowed = {os.path.join(ROOT, "tests", "golden", "make_golden.py")}
    for top in (PRODUCT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        for path in _sources(top):
            if path in allowed or path.endswith("test_product_isolation.py"):
                continue
            text = open(path).read()
            code = "\n".join(l for l in text.splitlines() if "/root/reference" in l and "open(" in l or "PYTHONPATH" in l)
            assert "/root/reference" not in code, f"{path} reads /root/reference at run time"
    bench = open(os.path.join(ROOT, "bench.py")).read()
    assert "/root/reference" not in bench


def test_kernels_have_no_host_side_twin():
    """The Python package holds no implementation of the game rules or of PUCT: those names only
    appear in csrc/."""
    for path in _sources(os.path.join(PRODUCT, "az_b200"), exts=(".py",)):
        text = open(path).read()
        assert "def select_leaf" not in text and "def backup" not in text and "def expand" not in text
