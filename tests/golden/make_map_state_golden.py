"""Pins oracle/map_state.py to the reference's own ``Agent_State`` bookkeeping methods and writes the fixture.

    python tests/golden/make_map_state_golden.py        (build container only: reads /root/reference)

nav/agent/agent_state.py cannot be imported here (skimage, skfmm, habitat are absent), so the UNMODIFIED sources of
``get_local_map_boundaries``, ``init_map_and_pose``, ``init_with_obs``, ``update_local_map`` and ``update_full_map`` are cut out of the file
with ``ast`` and attached to a stub class that provides exactly the attributes they read (args, CPU torch maps and poses,
planner_pose_inputs, selem_idx, global_goals, and a scripted ``sem_map_module``).  The stub follows the oracle's scripted
episodes (oracle.map_state.trajectory); after EVERY call all maps, poses, boundaries and scalars must be bit-equal.  The
fixture stores a digest of the state after every event.  Where the reference itself raises (explored disk past the high
edge of the local map: IndexError) the episode is truncated there and the fixture records the shorter length.
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
from oracle import map_state as O  # noqa: E402

SRC = "/root/reference/nav/agent/agent_state.py"
METHODS = ["get_local_map_boundaries", "init_map_and_pose", "init_with_obs", "update_local_map", "update_full_map"]


def reference_class():
    tree = ast.parse(open(SRC).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Agent_State")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in METHODS]
    assert len(fns) == len(METHODS)
    mod = ast.Module(body=fns, type_ignores=[])
    ns = {"np": np, "torch": torch}
    exec(compile(mod, SRC, "exec"), ns)
    return type("RefState", (), {m: ns[m] for m in METHODS})


def make_stub(Ref, o):
    s = Ref()
    s.args = types.SimpleNamespace(map_size_cm=o.map_size_cm, map_resolution=o.map_resolution,
                                   global_downscaling=o.global_downscaling, grid_resolution=o.grid_resolution,
                                   col_rad=o.col_rad, goal_reached_dist=o.goal_reached_dist)
    s.device = torch.device("cpu")
    s.full_w, s.full_h, s.local_w, s.local_h = o.full_w, o.full_h, o.local_w, o.local_h
    s.full_map = torch.zeros(o.nc, o.full_w, o.full_h)
    s.local_map = torch.zeros(o.nc, o.local_w, o.local_h)
    s.full_pose = torch.zeros(3)
    s.local_pose = torch.zeros(3)
    s.origins = np.zeros(3)
    s.lmb = np.zeros(4).astype(int)
    s.planner_pose_inputs = np.zeros(7)
    s.selem_idx = O.disk_idx(o.col_rad + 1)
    s.global_goals = [[0, 0]]
    s.poses = None
    s.loc_r = s.loc_c = 0
    s.dist_to_goal = 0.0
    return s


def check(tag, s, o):
    assert np.array_equal(s.full_map.numpy(), o.full_map), tag + ": full_map"
    assert np.array_equal(s.local_map.numpy(), o.local_map), tag + ": local_map"
    assert np.array_equal(s.full_pose.numpy(), o.full_pose), tag + ": full_pose"
    assert np.array_equal(s.local_pose.numpy(), o.local_pose), tag + ": local_pose"
    assert np.array_equal(np.asarray(s.origins), o.origins), tag + ": origins"
    assert [int(v) for v in s.lmb] == [int(v) for v in o.lmb], tag + ": lmb"
    assert np.array_equal(s.planner_pose_inputs, o.planner_pose_inputs), tag + ": planner_pose_inputs"
    if not tag.endswith("init"):
        assert s.loc_r == o.loc_r and s.loc_c == o.loc_c, tag + ": loc"
        assert float(s.dist_to_goal) == float(o.dist_to_goal), tag + ": dist_to_goal"


def run_case(Ref, case, out):
    name = case[0]
    s = None
    digests = []
    for i, (event, payload, o) in enumerate(O.trajectory(case)):
        if event == "init":
            s = make_stub(Ref, o)
            s.init_map_and_pose()
        elif event == "shift":
            s.local_pose = s.local_pose + torch.from_numpy(payload)
            s.update_full_map()
        elif event == "local":
            lm, pose, goal = payload
            s.global_goals = [list(goal)]
            s.sem_map_module = lambda obs, poses, lmap, lpose, st: (None, torch.from_numpy(lm.copy()), None,
                                                                    torch.from_numpy(pose.copy()))
            try:
                s.update_local_map(None)
            except IndexError:
                print(f"  {name}: the reference raised IndexError at event {i}; episode truncated")
                break
        else:
            s.update_full_map()
        check(f"{name} event {i} {event}", s, o)
        digests.append(O.digest(o))
    out[f"{name}_digests"] = np.stack(digests)
    print(f"{name}: {len(digests)} events, reference == oracle after every call; final lmb {o.lmb}, "
          f"explored cells {int(o.local_map[1].sum())}")


def check_init_with_obs(Ref, out):
    """init_with_obs (:103-146): first mapper call + 3x3 stamp at the cell of the local pose, incl. poses at the low edge."""
    rows = []
    for k in range(len(O.INIT_POSES)):
        o = O.init_with_obs_case(k)
        s = make_stub(Ref, o)
        s.init_map_and_pose()
        s.args.visualize = False
        lm, pose = o.local_map.copy(), o.local_pose.copy()
        s.sem_map_module = lambda obs, poses, lmap, lpose, st: (None, torch.from_numpy(lm.copy()), None,
                                                                torch.from_numpy(pose.copy()))
        s.init_with_obs(None, {"sensor_pose": [0., 0., 0.]})
        o.stamp_initial()
        assert np.array_equal(s.local_map.numpy(), o.local_map), f"init_with_obs pose {k}"
        rows.append(O.digest(o))
    out["init_with_obs_digests"] = np.stack(rows)
    print(f"init_with_obs: {len(rows)} poses, reference == oracle")


def main():
    Ref = reference_class()
    out = {}
    check_init_with_obs(Ref, out)
    for case in O.CASES:
        run_case(Ref, case, out)
    np.savez_compressed(os.path.join(HERE, "map_state.npz"), **out)


if __name__ == "__main__":
    main()
