"""CPU: the update_global_goal oracle (oracle/global_goal.py + oracle/fmm.c).  scikit-fmm (pinned at 2019.1.30 by the
reference, peanut.Dockerfile:8) is not installable here, so the fast-marching restatement has no reference-held vector
(parity unpinned); it is held to what any correct second-order fast marching satisfies - closed-form distances in free
space, the eikonal property |grad d| = 1, causality, detours around walls - and to the reference's own post-processing."""
import numpy as np
import numpy.ma as ma
import pytest

from oracle import global_goal as G


def scene(seed, n=240, rects=30):
    """Random obstacle rectangles on an n x n map, the source region kept free."""
    rng = np.random.default_rng(seed)
    m = np.zeros((n, n), np.float32)
    for _ in range(rects):
        r, c = rng.integers(0, n - 8, 2)
        h, w = rng.integers(2, n // 6, 2)
        m[r:r + h, c:c + w] = rng.choice([1.0, 0.6, 0.51])
    m[rng.integers(0, n, 200), rng.integers(0, n, 200)] = 0.5   # rint(0.5) == 0: not an obstacle
    return m


def test_free_space_is_euclidean_to_second_order():
    n, s = 201, 100
    dd = G.distance_field(np.ones((n, n), bool), s, s)
    yy, xx = np.mgrid[0:n, 0:n]
    eu = np.hypot(yy - s, xx - s)
    far = eu > 5
    fin = np.isfinite(dd)
    assert (~fin).sum() <= 4          # nothing is masked: the reference's post-processing turns the farthest cells into inf
    # up and left of the source the field is the plain second-order one (error < 0.3 cells over a radius of 140) ...
    q = far & fin & (yy <= s) & (xx <= s)
    assert np.abs(dd - eu)[q].max() < 0.3
    # ... right of / below it it carries the marcher's source artefact: the neighbour that freezes after its opposite one
    # takes a second-order step across the source (1/3 instead of 1), which shifts that half plane by at most one cell
    q = far & fin
    assert np.abs(dd - eu)[q].max() < 1.05
    assert dd[s, s] == 0 and dd[s - 1, s] == 1 and dd[s, s - 1] == 1
    assert abs(dd[s + 1, s] - 1 / 3) < 1e-12 and abs(dd[s, s + 1] - 1 / 3) < 1e-12


def test_eikonal_property_causality_and_masks():
    m = scene(1)
    trav = G.traversible(m, G.disk(4), np.zeros_like(m), np.zeros_like(m))
    src = (120, 120)
    trav[112:129, 112:129] = True
    dd = G.distance_field(trav, *src)
    assert np.isinf(dd[~trav]).all()
    reached = np.isfinite(dd)
    assert reached.sum() > 0.3 * dd.size
    # causality: every reached cell but the source has a 4-neighbour with a strictly smaller value
    pad = np.pad(dd, 1, constant_values=np.inf)
    nmin = np.minimum(np.minimum(pad[:-2, 1:-1], pad[2:, 1:-1]), np.minimum(pad[1:-1, :-2], pad[1:-1, 2:]))
    chk = reached.copy()
    chk[src] = False
    assert (nmin[chk] < dd[chk]).all()
    # |grad d| = 1 with the upwind first-order differences the scheme is built on, away from the source (loose: the field
    # itself is second order, and obstacles corners shed kinks)
    with np.errstate(invalid="ignore"):
        gy = dd - np.minimum(pad[:-2, 1:-1], pad[2:, 1:-1])
        gx = dd - np.minimum(pad[1:-1, :-2], pad[1:-1, 2:])
    g = np.sqrt(np.maximum(gy, 0) ** 2 + np.maximum(gx, 0) ** 2)
    sel = reached & (dd > 6)
    assert abs(np.median(g[sel]) - 1.0) < 0.02
    assert np.percentile(np.abs(g[sel] - 1.0), 99) < 0.35
    # geodesic >= straight line, everywhere
    yy, xx = np.mgrid[0:dd.shape[0], 0:dd.shape[1]]
    assert (dd[reached] >= np.hypot(yy - src[0], xx - src[1])[reached] - 1.05).all()


def test_wall_detour():
    n = 121
    trav = np.ones((n, n), bool)
    trav[20:101, 70] = False          # a wall; the shortest path to the far side goes around an end
    dd = G.distance_field(trav, 60, 50)
    direct = dd[60, 69]
    around = dd[60, 71]
    assert abs(direct - 19) < 1.1
    detour = np.hypot(60 - 19.5, 70 - 50) + np.hypot(60 - 19.5, 1)   # via the upper end (the lower one is symmetric)
    assert abs(around - detour) < 2.5 and around > direct + 40


def _fields(seed, n=240):
    rng = np.random.default_rng(100 + seed)
    m = scene(seed, n)
    coll = (rng.random((n, n)) < 0.002).astype(np.uint8)
    vis = np.zeros((n, n), np.uint8)
    vis[n // 2 - 3:n // 2 + 4, n // 4:3 * n // 4] = 1
    lmb = (n // 4, n // 4 + n // 2, n // 4, n // 4 + n // 2)
    loc = (n // 4, n // 4 + 3)
    tp = rng.random((n // 2, n // 2)).astype(np.float32) ** 4
    return m, coll, vis, lmb, loc, tp


def test_update_global_goal_flow():
    m, coll, vis, lmb, loc, tp = _fields(3)
    r = G.update_global_goal(m, coll, vis, lmb, loc[0], loc[1], tp.astype(np.float64), global_goals=[[5, 6]], last_global_goal=None)
    assert r["global_goals"] == [np.unravel_index(r["value"].argmax(), r["value"].shape)] and r["last_global_goal"] == [[5, 6]]
    assert r["dd_wt"].shape == tp.shape and 0 < r["dd_wt"].max() <= 1.0
    # same inputs again: the new goal equals the CURRENT goal, not the last one -> it is taken again (and last <- current)
    r2 = G.update_global_goal(m, coll, vis, lmb, loc[0], loc[1], tp.astype(np.float64), prev_dd_wt=r["dd_wt"],
                              global_goals=r["global_goals"], last_global_goal=r["last_global_goal"])
    assert r2["global_goals"] == r["global_goals"] and r2["last_global_goal"] == r["global_goals"]
    # ... and a third time it equals the last goal: suppressed, nothing changes
    r3 = G.update_global_goal(m, coll, vis, lmb, loc[0], loc[1], tp.astype(np.float64), prev_dd_wt=r2["dd_wt"],
                              global_goals=[(1, 1)], last_global_goal=r2["last_global_goal"])
    assert r3["global_goals"] == [(1, 1)] and r3["last_global_goal"] == r2["last_global_goal"]


def test_stuck_inside_obstacle_keeps_last_weights():
    n = 240
    m = np.zeros((n, n), np.float32)
    m[50:70, 50:70] = 1.0                      # the agent sits inside a block (after dilation nothing around it is free)
    lmb = (0, 120, 0, 120)
    tp = np.ones((120, 120))
    prev = np.full((120, 120), 0.25)
    r = G.update_global_goal(m, np.zeros_like(m), np.zeros_like(m), lmb, 60, 60, tp, prev_dd_wt=prev, global_goals=[[0, 0]])
    assert np.isfinite(r["dd"]).sum() == 1 and r["dd_wt"] is prev
    r0 = G.update_global_goal(m, np.zeros_like(m), np.zeros_like(m), lmb, 60, 60, tp, prev_dd_wt=None, global_goals=[[0, 0]])
    assert r0["dd_wt"].sum() == 1.0 and r0["global_goals"] == [(60, 60)]


@pytest.mark.parametrize("temp", [-1.0, 0.0])
def test_special_temperatures(temp):
    m, coll, vis, lmb, loc, tp = _fields(5)
    with np.errstate(divide="ignore"):
        r = G.update_global_goal(m, coll, vis, lmb, loc[0], loc[1], tp.astype(np.float64), dist_weight_temperature=temp,
                                 global_goals=[[0, 0]])
    if temp == -1.0:
        assert r["value"] is not None and np.array_equal(r["value"], tp.astype(np.float64))
    else:
        w = r["value"]
        assert w.max() <= np.exp(-60 / 100.) + 1e-12
