"""Bit-reproducibility across processes (the ranks of one job must agree with each other): different conv launch
configurations differ in fp32 summation order, and a configuration picked by timing can differ from process to process.
Processes that import the same launch-configuration table (peanut_b200/tuning/*.txt, $PN_CONV_TUNING_FILE, or
parallel.build_synchronised's broadcast) must therefore produce identical bits, and must not time anything."""
import hashlib
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import hashlib, os, sys
sys.path.insert(0, %r)
import numpy as np, torch
from oracle import prednet as OC
from peanut_b200 import _lib, prediction as P
seg = P.Segmentor(P._default_cfg(24, 6), OC.synth_state_dict(24, 6, seed=0), "cuda:0", precision="bf16")
x = torch.from_numpy(np.stack([OC.synth_partial_map(24, 64, 64, seed=i) for i in range(2)])).cuda()
y = seg.forward_device(x, apply_sigmoid=False)
y = seg.forward_device(x, apply_sigmoid=False)   # second call = CUDA graph replay
torch.cuda.synchronize()
print("HASH", hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest())
out = os.environ.get("PN_TEST_TUNING_OUT")
if out:
    open(out, "w").write(_lib.tuning_export())
""" % ROOT


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    h = [l.split()[1] for l in r.stdout.splitlines() if l.startswith("HASH")]
    assert len(h) == 1
    return h[0], r.stderr


def test_two_processes_bit_equal_with_shared_table(tmp_path):
    table = str(tmp_path / "tuning.txt")
    # process 1 tunes by timing (committed tables ignored so that it really has to) and exports what it chose
    h1, _ = _run({"PN_CONV_TUNING_FILE": "none", "PN_TEST_TUNING_OUT": table})
    # ... but "none" also skips the import of the override, so hand the table over explicitly to processes 2 and 3
    text = open(table).read()
    assert len(text.splitlines()) > 10
    h2, err2 = _run({"PN_CONV_TUNING_FILE": table, "PN_CONV_TUNE_LOG": "1"})
    h3, err3 = _run({"PN_CONV_TUNING_FILE": table, "PN_CONV_TUNE_LOG": "1"})
    assert h1 == h2 == h3
    assert "[tune]" not in err2 and "[tune]" not in err3   # nothing was timed: every layer came from the table


def test_table_mode_never_times():
    h1, err1 = _run({"PN_CONV_TUNING_FILE": "none", "PN_CONV_AUTOTUNE": "table", "PN_CONV_TUNE_LOG": "1"})
    h2, err2 = _run({"PN_CONV_TUNING_FILE": "none", "PN_CONV_AUTOTUNE": "table", "PN_CONV_TUNE_LOG": "1"})
    assert h1 == h2 and "[tune]" not in err1 + err2
