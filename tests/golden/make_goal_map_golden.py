"""Pins oracle/goal_map.py to the reference's own ``Agent_State.update_goal_map`` and writes tests/golden/goal_map.npz.

    python tests/golden/make_goal_map_golden.py        (build container only: reads /root/reference)

The UNMODIFIED method source is cut out of nav/agent/agent_state.py with ``ast`` and run on a stub state (CPU torch
local_map, args.only_explore / goal_erode, global_goals, goal_cat).  scikit-image is absent here, so the name ``skimage`` the
method refers to is bound to a two-function stand-in restated from scikit-image's published wrappers (see
oracle/goal_map.py); scipy.ndimage does the morphology.  goal_map must match in values and dtype, found_goal in value.
"""
import ast
import os
import sys
import types

import numpy as np
import torch
from scipy import ndimage as ndi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import goal_map as O  # noqa: E402

SRC = "/root/reference/nav/agent/agent_state.py"


def skimage_standin():
    def default_footprint(image):
        return ndi.generate_binary_structure(image.ndim, 1)

    def binary_erosion(image, footprint=None, out=None):
        if out is None:
            out = np.empty(image.shape, dtype=bool)
        ndi.binary_erosion(image, structure=default_footprint(image) if footprint is None else footprint, output=out,
                           border_value=True)
        return out

    def binary_dilation(image, footprint=None, out=None):
        if out is None:
            out = np.empty(image.shape, dtype=bool)
        ndi.binary_dilation(image, structure=default_footprint(image) if footprint is None else footprint, output=out)
        return out

    return types.SimpleNamespace(morphology=types.SimpleNamespace(binary_erosion=binary_erosion, binary_dilation=binary_dilation))


def reference_method():
    tree = ast.parse(open(SRC).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Agent_State")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "update_goal_map")
    ns = {"np": np, "torch": torch, "skimage": skimage_standin()}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), ns)
    return ns["update_goal_map"]


# (seed, nc, n, goal_cat, goal_name, goal_erode, only_explore, blank goal channel)
CASES = [
    (1, 14, 96, 1, "chair", 3, 0, False),
    (2, 14, 96, 5, "tv_monitor", 3, 0, False),     # 'tv' in the name: no erosion / dilation, float32 result
    (3, 14, 120, 0, "bed", 1, 0, False),
    (4, 14, 96, 3, "toilet", 0, 0, False),         # goal_erode 0: dilation only
    (5, 14, 96, 2, "sofa", 3, 0, True),            # goal never seen -> goal_map = the long-term goal cell
    (6, 14, 96, 4, "plant", 3, 1, False),          # only_explore
    (7, 14, 96, 1, "chair", 6, 0, False),          # everything eroded away -> not found
    (8, 12, 64, 7, "cup", 2, 0, False),            # goal channel outside 4:10 (never cancels in the sum)
]


def main():
    ref_fn = reference_method()
    out = {}
    for i, (seed, nc, n, goal_cat, name, erode, only_explore, blank) in enumerate(CASES):
        lm = O.synth_local_map(seed, nc, n, goal_cat=goal_cat)
        if blank:
            lm[goal_cat + 4] = 0
        goal = [int(seed * 7 % n), int(seed * 13 % n)]
        stub = types.SimpleNamespace()
        stub.args = types.SimpleNamespace(only_explore=only_explore, goal_erode=erode)
        stub.local_w, stub.local_h = n, n
        stub.local_map = torch.from_numpy(lm.copy())
        stub.global_goals = [list(goal)]
        stub.goal_cat = goal_cat
        ref_fn(stub, {"goal_name": name})
        gm, found = O.update_goal_map(lm, goal_cat, goal, name, erode, only_explore)
        assert found == stub.found_goal, (i, found, stub.found_goal)
        assert gm.dtype == stub.goal_map.dtype, (i, gm.dtype, stub.goal_map.dtype)
        assert np.array_equal(gm, stub.goal_map), f"case {i}: goal_map differs"
        out[f"goal_map_{i}"] = np.packbits(stub.goal_map != 0)
        out[f"meta_{i}"] = np.array([seed, nc, n, goal_cat, erode, only_explore, int(blank), found, goal[0], goal[1],
                                     int(stub.goal_map.dtype == np.float64)], np.int64)
        out[f"name_{i}"] = np.array(name)
        print(f"case {i} ({name}, erode {erode}): reference == oracle, found_goal {found}, {int((gm != 0).sum())} goal cells, "
              f"dtype {gm.dtype}")
    np.savez_compressed(os.path.join(HERE, "goal_map.npz"), **out)


if __name__ == "__main__":
    main()
