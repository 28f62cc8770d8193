"""Pins oracle/agent_prediction.py to the reference's own ``Agent_State.update_prediction`` and writes the fixture.

    python tests/golden/make_update_prediction_golden.py        (build container only: reads /root/reference)

nav/agent/agent_state.py cannot be imported here (skimage, skfmm, habitat are absent), so the UNMODIFIED source of the one
method is cut out of the file with ``ast`` and executed on a stub object that provides exactly the attributes it reads
(args.prediction_window, full_map / local_map as CPU torch tensors, lmb, full_w / full_h, goal_cat, prediction_model).
Bit-equality of target_pred (values AND dtype) and of the updated full_map with the oracle is required for every case.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import agent_prediction as O  # noqa: E402

SRC = "/root/reference/nav/agent/agent_state.py"


def reference_method():
    tree = ast.parse(open(SRC).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Agent_State")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "update_prediction")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np, "torch": torch}
    exec(compile(mod, SRC, "exec"), ns)
    return ns["update_prediction"]


def main():
    ref_fn = reference_method()
    out = {}
    # (full, local, window, seed, goal): crop branch, full-window branch, window at an edge of the local map
    cases = [(96, 48, 72, 1, 2), (96, 48, 96, 2, 0), (120, 60, 72, 3, 5), (96, 48, 40, 4, 3)]
    for i, (full, local, win, seed, goal) in enumerate(cases):
        fm, lm, lmb = O.synth_state(seed, full, local)
        stub = types.SimpleNamespace()
        stub.args = types.SimpleNamespace(prediction_window=win)
        stub.full_map = torch.from_numpy(fm.copy())
        stub.local_map = torch.from_numpy(lm.copy())
        stub.lmb = lmb.copy()
        stub.full_w, stub.full_h = full, full
        stub.goal_cat = goal
        stub.prediction_model = types.SimpleNamespace(get_prediction=O.fake_prediction)
        ref_fn(stub)
        fm_o = fm.copy()
        tp = O.update_prediction(fm_o, lm, lmb, goal, O.fake_prediction, win)
        assert tp.dtype == stub.target_pred.dtype, (tp.dtype, stub.target_pred.dtype)
        assert np.array_equal(tp, stub.target_pred), f"case {i}: target_pred differs"
        assert np.array_equal(fm_o, stub.full_map.numpy()), f"case {i}: full_map differs"
        out[f"target_pred_{i}"] = stub.target_pred
        out[f"meta_{i}"] = np.array([full, local, win, seed, goal], np.int64)
        print(f"case {i}: full {full} local {local} window {win}: reference == oracle, dtype {tp.dtype}, "
              f"{int((tp > 0).sum())} non-zero cells")
    np.savez_compressed(os.path.join(HERE, "update_prediction.npz"), **out)


if __name__ == "__main__":
    main()
