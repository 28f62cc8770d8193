"""GPU: the device-resident map bookkeeping (pn_map_init / pn_map_stamp_initial / pn_map_update_local /
pn_map_update_full through peanut_b200.map_state.MapState) must be BIT-EXACT against oracle/map_state.py, which is pinned to
the unmodified Agent_State methods (nav/agent/agent_state.py:153-211, 116-122, 276-303, 308-338), after every call of the
scripted episodes - maps, poses, window bounds, origins, planner pose vector, agent cell and distance to goal."""
import numpy as np
import pytest
import torch

from oracle import map_state as O
from peanut_b200.map_state import MapState

pytestmark = pytest.mark.gpu


def _device_state(ctx, case, E=1, f64_cells=False):
    name, nc, size_cm, res, gds, grid, col_rad, *_ = case
    return MapState(ctx, E, num_sem_categories=nc - 4, map_size_cm=size_cm, map_resolution=res, global_downscaling=gds,
                    grid_resolution=grid, col_rad=col_rad, goal_reached_dist=75.0, f64_cells=f64_cells)


def _compare(tag, d, e, o, small=True):
    assert np.array_equal(d.full_map[e].cpu().numpy(), o.full_map), tag + ": full_map"
    assert np.array_equal(d.local_map[e].cpu().numpy(), o.local_map), tag + ": local_map"
    assert np.array_equal(d.full_pose[e].cpu().numpy(), o.full_pose), tag + ": full_pose"
    assert np.array_equal(d.local_pose[e].cpu().numpy(), o.local_pose), tag + ": local_pose"
    assert np.array_equal(d.origins[e].cpu().numpy(), o.origins), tag + ": origins"
    assert d.lmb[e].tolist() == [int(v) for v in o.lmb], tag + ": lmb"
    assert np.array_equal(d.planner_pose_inputs[e].cpu().numpy(), o.planner_pose_inputs), tag + ": planner_pose_inputs"
    if small:
        assert d.loc[e].tolist() == [o.loc_r, o.loc_c], tag + ": loc"
        assert float(d.dist_to_goal[e]) == float(o.dist_to_goal), tag + ": dist_to_goal"


def _apply(d, e_slice, event, payload):
    """Apply one oracle event to environment(s) of the device state (the other environments see the same calls)."""
    if event == "init":
        d.init_map_and_pose()
    elif event == "shift":
        d.local_pose[e_slice] += torch.from_numpy(payload).cuda()
        d.update_full_map()
    elif event == "local":
        lm, pose, goal = payload
        d.local_map[e_slice] = torch.from_numpy(lm).cuda()
        d.local_pose[e_slice] = torch.from_numpy(pose).cuda()
        d.global_goals[e_slice] = torch.tensor(goal, dtype=torch.int32).cuda()
        d.update_local_map()
    else:
        d.update_full_map()


@pytest.mark.parametrize("f64_cells", [False, True], ids=["f32cells", "f64cells"])
@pytest.mark.parametrize("case", O.CASES, ids=[c[0] for c in O.CASES])
def test_episode_bit_exact(ctx, case, f64_cells):
    d = _device_state(ctx, case, 1, f64_cells)
    for i, (event, payload, o) in enumerate(O.trajectory(case, f64_cells=f64_cells)):
        _apply(d, slice(0, 1), event, payload)
        _compare(f"{case[0]} event {i} {event}", d, 0, o, small=event != "init")


def test_batched_environments_are_independent(ctx):
    """Three environments walk three different episodes of the same geometry inside one MapState (same seeds as the
    fixtures, different step sizes): every call updates all of them, each must equal its own oracle."""
    base = O.CASES[1]
    cases = [base, base[:9] + (4.0, base[10], 11), base[:7] + ((30, -40),) + base[8:11] + (12,)]
    d = _device_state(ctx, base, len(cases))
    gens = [O.trajectory(c) for c in cases]
    for i, events in enumerate(zip(*gens)):
        kinds = {ev[0] for ev in events}
        assert len(kinds) == 1  # same schedule (steps, num_local_steps) for all three
        kind = kinds.pop()
        if kind == "init":
            d.init_map_and_pose()
        elif kind == "shift":
            for e, ev in enumerate(events):
                d.local_pose[e] += torch.from_numpy(ev[1]).cuda()
            d.update_full_map()
        elif kind == "local":
            for e, ev in enumerate(events):
                lm, pose, goal = ev[1]
                d.local_map[e] = torch.from_numpy(lm).cuda()
                d.local_pose[e] = torch.from_numpy(pose).cuda()
                d.global_goals[e] = torch.tensor(goal, dtype=torch.int32).cuda()
            d.update_local_map()
        else:
            d.update_full_map()
        for e, ev in enumerate(events):
            _compare(f"env {e} event {i} {kind}", d, e, ev[2], small=kind != "init")


def test_init_with_obs_stamp(ctx):
    for k in range(len(O.INIT_POSES)):
        o = O.init_with_obs_case(k)
        d = _device_state(ctx, ("x", 5, 960, 5, 2, 24, 4), 1)
        d.init_map_and_pose()
        d.local_map[0] = torch.from_numpy(o.local_map).cuda()
        d.local_pose[0] = torch.from_numpy(o.local_pose).cuda()
        d.stamp_initial()
        o.stamp_initial()
        assert np.array_equal(d.local_map[0].cpu().numpy(), o.local_map), f"pose {k}"


def test_reference_geometry_roundtrip(ctx):
    """Full-size maps (4800 cm / 5 cm -> 960^2, local 480^2, 14 channels, 2 environments): store-back + recentre + re-cut
    keeps every cell (checksum of the full map is the checksum of its parts) and the window follows the agent."""
    d = MapState(ctx, 2)
    d.init_map_and_pose()
    assert d.lmb.tolist() == [[240, 720, 240, 720]] * 2
    g = torch.Generator(device="cuda").manual_seed(3)
    d.local_map.copy_(torch.rand(d.local_map.shape, generator=g, device="cuda"))
    before = d.local_map.clone()
    d.local_pose[0, :2] += torch.tensor([3.0, -2.0], device="cuda")  # 60 cells right, 40 cells up: window moves by (-48, +48)
    d.update_full_map()
    torch.cuda.synchronize()
    assert d.lmb.tolist() == [[192, 672, 288, 768], [240, 720, 240, 720]]
    assert torch.equal(d.full_map[:, :, 240:720, 240:720], before)            # written back at the OLD window
    assert torch.equal(d.local_map[0], d.full_map[0, :, 192:672, 288:768])    # re-cut at the NEW window
    assert torch.equal(d.local_map[1], before[1])
    info = d.planner_inputs()
    assert info["lmb"].tolist() == d.lmb.tolist() and info["pose_pred"].shape == (2, 7)
    assert d.full_pose[0].tolist() == [27.0, 22.0, 0.0]


def test_missing_goal_pointer_is_an_error(ctx):
    import ctypes
    from peanut_b200 import _lib
    d = MapState(ctx, 1, map_size_cm=480)
    arrays = d._arrays()
    arrays.global_goal = None
    rc = ctx.lib.pn_map_update_local(ctx.handle, ctypes.byref(d.cfg), ctypes.byref(arrays), 1, None)
    assert rc != 0 and b"global_goal" in ctx.lib.pn_last_error()
