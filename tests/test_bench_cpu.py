"""CPU: the parts of bench.py's contract that do not need a GPU - workload resolution (fixed per-GPU work at every N), the
workload description both arms print verbatim, and the keys the JSON lines must carry."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("pn_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_same_per_gpu_workload_at_every_world_size():
    b = _bench()
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = b.parse()
    finally:
        sys.argv = argv
    assert a.gpus == 1 and a.warmup >= 3 and a.workload == "auto" and a.mode == "dependent"
    names = {w: b.resolve_workload(a, w) for w in (1, 2, 4, 8)}
    assert set(names.values()) == {"cfg2"}                      # weak scaling: per-GPU work fixed as N grows
    assert b.WORKLOADS["cfg2"]["envs"] == 32 and b.WORKLOADS["cfg3"]["envs"] == 8 and b.WORKLOADS["cfg1"]["envs"] == 1
    assert b.WORKLOADS["cfg1"]["precision"] == "tf32" and b.WORKLOADS["cfg2"]["precision"] == "bf16"
    assert b.MAP_SHAPES["base"] == (24, 240, 240)                # BASELINE.json: 24 x 240 x 240 partial maps


def test_both_arms_describe_the_workload_identically():
    b = _bench()
    wl = b.WORKLOADS["cfg2"]
    c1 = b.shared_config(wl["desc"], wl["envs"], b.MAP_SHAPES["base"], "dependent")
    c2 = b.shared_config(wl["desc"], wl["envs"], b.MAP_SHAPES["base"], "dependent")
    assert c1 == c2 and c1["frame"] == [480, 640] and c1["map_shape"] == [24, 240, 240] and c1["envs_per_gpu"] == 32
    assert "flush" in c1["l2"] and "max over ranks" in c1["timing"]      # the timing rules the contract asks `config` to state
    assert "reference order" in c1["mode"]
    # both arms build `config` through this one function
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": shared_config(') == 2
