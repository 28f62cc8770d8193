"""CPU: oracle/goal_map.py against the fixture written while the UNMODIFIED Agent_State.update_goal_map
(nav/agent/agent_state.py:423-452) agreed with it (tests/golden/make_goal_map_golden.py)."""
import os

import numpy as np
import pytest

from oracle import goal_map as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "goal_map.npz"))
N_CASES = sum(1 for k in GOLD.files if k.startswith("meta_"))


def golden_case(i):
    seed, nc, n, goal_cat, erode, only_explore, blank, found, g0, g1, is_f64 = (int(v) for v in GOLD[f"meta_{i}"])
    lm = O.synth_local_map(seed, nc, n, goal_cat=goal_cat)
    if blank:
        lm[goal_cat + 4] = 0
    want = np.unpackbits(GOLD[f"goal_map_{i}"])[:n * n].reshape(n, n).astype(bool)
    return lm, goal_cat, [g0, g1], str(GOLD[f"name_{i}"]), erode, only_explore, want, found, is_f64


@pytest.mark.parametrize("i", range(N_CASES))
def test_oracle_matches_reference_fixture(i):
    lm, goal_cat, goal, name, erode, only_explore, want, found, is_f64 = golden_case(i)
    gm, f = O.update_goal_map(lm, goal_cat, goal, name, erode, only_explore)
    assert f == found
    assert (gm.dtype == np.float64) == bool(is_f64)
    assert np.array_equal(gm != 0, want)
    assert set(np.unique(gm)) <= {0.0, 1.0}


def test_n_cross_erosions_equal_one_diamond_erosion():
    """The device kernel erodes once with the diamond |dr| + |dc| <= n instead of n times with the cross."""
    from scipy import ndimage as ndi
    rng = np.random.default_rng(0)
    img = ndi.binary_dilation(rng.random((80, 90)) < 0.02, iterations=6)
    for n in (1, 2, 3, 5):
        it = img.copy()
        for _ in range(n):
            it = ndi.binary_erosion(it, structure=O.CROSS, border_value=True)
        ax = np.arange(-n, n + 1)
        diamond = (np.abs(ax)[:, None] + np.abs(ax)[None, :]) <= n
        assert np.array_equal(it, ndi.binary_erosion(img, structure=diamond, border_value=True))
