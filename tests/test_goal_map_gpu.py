"""GPU: pn_goal_map (through MapState.update_goal_map) must be BIT-EXACT against oracle/goal_map.py, which is pinned to the
unmodified Agent_State.update_goal_map (nav/agent/agent_state.py:423-452): same goal cells, same found_goal."""
import numpy as np
import pytest
import torch

from oracle import goal_map as O
from peanut_b200.map_state import MapState
from tests.test_goal_map_cpu import N_CASES, golden_case

pytestmark = pytest.mark.gpu


def _state(ctx, E, nc, n):
    d = MapState(ctx, E, num_sem_categories=nc - 4, map_size_cm=n * 5, map_resolution=5, global_downscaling=1)
    assert (d.local_w, d.local_h) == (n, n)
    return d


@pytest.mark.parametrize("i", range(N_CASES))
def test_golden_cases(ctx, i):
    lm, goal_cat, goal, name, erode, only_explore, want, found, _ = golden_case(i)
    d = _state(ctx, 1, lm.shape[0], lm.shape[1])
    d.local_map[0] = torch.from_numpy(lm).cuda()
    d.global_goals[0] = torch.tensor(goal, dtype=torch.int32).cuda()
    gm, f = d.update_goal_map([goal_cat], [name], goal_erode=erode, only_explore=only_explore)
    assert int(f[0]) == found
    got = gm[0].cpu().numpy()
    assert np.array_equal(got != 0, want)
    assert set(np.unique(got)) <= {0.0, 1.0}


def test_batched_random_maps(ctx):
    """8 environments, different goals / names / maps in one launch, reference-size local map (480 x 480), erosion 3 and
    the tile seams of the 32 x 32 blocks crossed by blobs; then the same state with erosion 0, 1 and 5."""
    E, nc, n = 8, 14, 480
    d = _state(ctx, E, nc, n)
    rng = np.random.default_rng(5)
    maps = []
    for e in range(E):
        m = np.zeros((nc, n, n), np.float32)
        for c in range(4, nc):
            from scipy import ndimage as ndi
            seeds = rng.random((n, n)) < 0.0004
            m[c] = ndi.binary_dilation(seeds, iterations=int(rng.integers(2, 9))) * rng.random((n, n)).astype(np.float32)
        if e == 3:
            m[4 + 3] = 0  # goal never seen
        maps.append(m)
    d.local_map.copy_(torch.from_numpy(np.stack(maps)))
    goal_cats = [e % 6 for e in range(E)]
    names = ["chair", "tv_monitor", "bed", "toilet", "sofa", "tv", "plant", "couch"]
    goals = rng.integers(0, n, (E, 2))
    d.global_goals.copy_(torch.from_numpy(goals.astype(np.int32)))
    for erode in (3, 0, 1, 5):
        gm, f = d.update_goal_map(goal_cats, names, goal_erode=erode)
        gm, f = gm.cpu().numpy(), f.cpu().numpy()
        for e in range(E):
            want, wf = O.update_goal_map(maps[e], goal_cats[e], goals[e], names[e], erode)
            assert int(f[e]) == wf, (erode, e)
            assert np.array_equal(gm[e], want.astype(np.float32)), (erode, e)
    assert not f[3] and gm[3].sum() == 1 and gm[3, goals[3][0], goals[3][1]] == 1


def test_only_explore(ctx):
    d = _state(ctx, 2, 14, 64)
    d.local_map.uniform_(0, 1)
    d.global_goals.copy_(torch.tensor([[3, 5], [60, 1]], dtype=torch.int32))
    gm, f = d.update_goal_map([1, 2], ["chair", "bed"], only_explore=1)
    assert f.tolist() == [0, 0] and gm.sum().item() == 2 and gm[0, 3, 5] == 1 and gm[1, 60, 1] == 1
