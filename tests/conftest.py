import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "custom-alphazero_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer CPU test")
    # The native artefacts are build products (git-ignored).  On a fresh checkout build them once, exactly as
    # __graft_entry__.build() does (nvcc cross-compiles sm_100a without a GPU); a failed build fails the tests loudly.
    lib = os.path.join(PKG, "libaz_b200.so")
    if not os.path.exists(lib):
        import subprocess

        subprocess.check_call(["sh", os.path.join(PKG, "build.sh")])


@pytest.fixture(scope="session")
def golden():
    import json

    def load(name):
        with open(os.path.join(ROOT, "tests", "golden", name + ".json")) as fp:
            return json.load(fp)

    return load
