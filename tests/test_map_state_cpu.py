"""CPU: oracle/map_state.py replays the scripted episodes and must reproduce the digests that
tests/golden/make_map_state_golden.py recorded while the UNMODIFIED reference methods (Agent_State.init_map_and_pose /
init_with_obs / update_local_map / update_full_map, nav/agent/agent_state.py) agreed with it bit for bit."""
import os

import numpy as np
import pytest

from oracle import map_state as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "map_state.npz"))


@pytest.mark.parametrize("case", O.CASES, ids=[c[0] for c in O.CASES])
def test_oracle_reproduces_reference_digests(case):
    want = GOLD[case[0] + "_digests"]
    got = []
    for event, payload, o in O.trajectory(case):
        got.append(O.digest(o))
        if len(got) == len(want):  # the reference itself stopped here (IndexError past the high edge) or the episode ended
            break
    assert len(got) == len(want)
    assert np.array_equal(np.stack(got), want)


def test_init_with_obs_digests():
    want = GOLD["init_with_obs_digests"]
    for k in range(len(O.INIT_POSES)):
        o = O.init_with_obs_case(k)
        o.stamp_initial()
        assert np.array_equal(O.digest(o), want[k])


def test_boundaries_properties():
    """Window always inside the full map, of the local size, origin on the grid unless clamped (agent_state.py:153-177)."""
    rng = np.random.default_rng(0)
    for _ in range(2000):
        full = int(rng.integers(40, 400)) * 2
        local = full // 2
        grid = int(rng.integers(1, 40))
        r, c = int(rng.integers(-20, full + 20)), int(rng.integers(-20, full + 20))
        x1, x2, y1, y2 = O.boundaries(r, c, local, local, full, full, 2, grid)
        assert 0 <= x1 and x2 <= full and x2 - x1 == local and 0 <= y1 and y2 <= full and y2 - y1 == local
        assert x1 % grid == 0 or x1 in (0, full - local)
    assert O.boundaries(5, 7, 50, 50, 50, 50, 1, 24) == [0, 50, 0, 50]


def test_disk_matches_skimage_definition():
    rr, cc = O.disk_idx(5)
    assert len(rr) == 81 and rr.min() == 0 and rr.max() == 10  # skimage.morphology.disk(5).sum() == 81
    m = np.zeros((11, 11), int)
    m[rr, cc] = 1
    assert np.array_equal(m, m.T) and m[5, 0] == 1 and m[0, 0] == 0 and m[1, 2] == 1 and m[1, 1] == 0
