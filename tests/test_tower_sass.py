"""The SASS of the net kernel (csrc/az_tower.cu) keeps the properties its speed depends on.  No GPU needed: cuobjdump
reads the sm_100a code out of the built library.

* tcgen05 tensor cores, bulk copies and TMEM loads are there (UTCHMMA / UBLKCP / LDTM), no legacy HMMA;
* no uniform-datapath instruction is wrapped in a broadcast loop (`BRA.U.ANY` after ELECT / R2UR.BROADCAST): that is what
  the compiler emits for `cp.async.bulk` / `tcgen05.mma` issued inside `if (lane == 0)`, and it made the weight producer
  take 378 cycles per stage where the tensor pipe needs one every 256 (DESIGN.md section 4,
  profiles/net_kernel_issue_study_r2.json)."""
import os
import re
import shutil
import subprocess

import pytest

from tests.helpers import ROOT

LIB = os.path.join(ROOT, "custom-alphazero_b200", "libaz_b200.so")


def _tower_functions():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(LIB):
        pytest.skip("libaz_b200.so not built")
    text = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", text)
    return {b.split("\n", 1)[0].strip(): b for b in blocks[1:] if "k_tower" in b.split("\n", 1)[0]}


def test_net_kernel_sass_keeps_the_uniform_datapath():
    funcs = _tower_functions()
    # <NET, PAIR, TREES>: tower alone and whole net, single CTA and CTA pair, and the whole net with tree warps
    assert len(funcs) >= 6, sorted(funcs)
    for name, sass in funcs.items():
        assert "UTCHMMA" in sass and "UBLKCP" in sass and "LDTM" in sass, name
        assert not re.search(r"\bHMMA\b", sass), f"{name}: legacy tensor-core instruction"
        assert "BRA.U.ANY" not in sass, f"{name}: a uniform-datapath instruction sits in a broadcast loop (issued from divergent code)"
        pair = "ILb0ELb1E" in name or "ILb1ELb1E" in name
        assert ("UTCHMMA.2CTA" in sass) == pair, name
        # the rolled filter-row loop: a handful of MMA sites, not 38 stages of straight-line code per block
        assert len(re.findall(r"UTCHMMA", sass)) <= 48, name
