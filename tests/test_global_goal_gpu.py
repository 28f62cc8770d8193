"""N1b on the GPU: MapState.update_global_goal (pn_global_goal) against the oracle (oracle/global_goal.py, whose fast
marching is a restatement of scikit-fmm 2019.1.30: parity unpinned, see that file).

  * traversible mask (dilation, collision / visited overrides, agent cell) and therefore the SET of reachable cells: exact;
  * geodesic distance: the marcher's first cells (distance <= 3) bit-exact - the device replays the marcher there -, the rest
    within 0.5 cells (max) / 1e-2 cells (mean): the device solves the marcher's discretisation as a fixed point, the
    marcher's own result depends on the order in which it froze cells where second-order stencils switch on and off
    (measured on B200: max 0.08 / mean 1e-3 on 240 x 240 maps, max 0.24 / mean 3.2e-3 on 960 x 960 maps with paths of
    ~980 cells, i.e. 2.4e-4 of the distance);
  * weights / value: 5e-3 (exp of the above over the temperature 100);
  * goal cell: equal whenever the oracle's best value beats every cell outside a 2-cell neighbourhood of it by more than
    1e-2 of itself; goal bookkeeping (last goal, kinds, "stuck" rule): exact.
"""
import numpy as np
import pytest
import torch

from oracle import global_goal as G
from peanut_b200 import _lib
from peanut_b200.map_state import MapState
from tests.test_global_goal_cpu import scene

pytestmark = pytest.mark.gpu


def make_state(ctx, E, n, seeds, device="cuda:0"):
    """MapState with full maps of n x n cells (local = n / 2), random obstacle scenes, windows and agent cells."""
    st = MapState(ctx, E, num_sem_categories=10, map_size_cm=n * 5, map_resolution=5, global_downscaling=2)
    host = dict(m=[], coll=[], vis=[], lmb=[], loc=[], tp=[])
    for e, seed in enumerate(seeds):
        rng = np.random.default_rng(1000 + seed)
        m = scene(seed, n, rects=int(10 * n / 240))   # sparse enough that most of the map stays connected after dilation
        coll = (rng.random((n, n)) < 0.002).astype(np.uint8)
        vis = np.zeros((n, n), np.uint8)
        r0 = int(rng.integers(0, n // 2 + 1)) // 4 * 4
        c0 = int(rng.integers(0, n // 2 + 1)) // 4 * 4
        lmb = (r0, r0 + n // 2, c0, c0 + n // 2)
        loc = (int(rng.integers(5, n // 2 - 5)), int(rng.integers(5, n // 2 - 5)))
        ar, ac = loc[0] + r0, loc[1] + c0
        vis[max(ar - 6, 0):ar + 7, max(ac - 6, 0):ac + 7] = 1      # the agent has been here: free whatever the map says
        tp = (rng.random((n // 2, n // 2)) ** 6).astype(np.float32)
        for k, v in zip(("m", "coll", "vis", "lmb", "loc", "tp"), (m, coll, vis, lmb, loc, tp)):
            host[k].append(v)
        st.full_map[e, 0] = torch.from_numpy(m)
    st.lmb.copy_(torch.tensor(host["lmb"], dtype=torch.int32))
    st.loc.copy_(torch.tensor(host["loc"], dtype=torch.int32))
    st.global_goals.copy_(torch.tensor([[3, 4]] * E, dtype=torch.int32))
    dev = dict(coll=torch.from_numpy(np.stack(host["coll"])).cuda(), vis=torch.from_numpy(np.stack(host["vis"])).cuda(),
               tp=torch.from_numpy(np.stack(host["tp"])).cuda())
    return st, host, dev


def oracle_env(host, e, **kw):
    return G.update_global_goal(host["m"][e], host["coll"][e], host["vis"][e], host["lmb"][e], host["loc"][e][0], host["loc"][e][1],
                                host["tp"][e].astype(np.float64), **kw)


@pytest.mark.parametrize("n,E", [(240, 3), (960, 2)])
def test_distance_field_and_goal(ctx, n, E):
    st, host, dev = make_state(ctx, E, n, seeds=list(range(10, 10 + E)))
    goals = st.update_global_goal(dev["tp"], 500.0, dev["coll"], dev["vis"])
    torch.cuda.synchronize()
    dd = st.dd.cpu().numpy()
    for e in range(E):
        ref = oracle_env(host, e, global_goals=[[3, 4]], last_global_goal=None)
        rd = ref["dd"]
        assert np.array_equal(np.isfinite(dd[e]), np.isfinite(rd)), "reachable sets differ"
        fin = np.isfinite(rd)
        assert fin.sum() > 0.1 * rd.size
        with np.errstate(invalid="ignore"):
            diff = np.abs(dd[e] - rd)[fin]
        assert diff.max() <= 0.5 and diff.mean() <= 1e-2, (diff.max(), diff.mean())
        near = fin & (rd <= 3.0)
        assert near.sum() >= 5 and np.array_equal(dd[e][near], rd[near])           # the replayed seed: bit-exact
        wt = st.dd_wt[e].cpu().numpy()
        assert np.abs(wt - ref["dd_wt"]).max() <= 5e-3
        val = st.value[e].cpu().numpy()
        assert np.abs(val - ref["value"]).max() <= 5e-3 * max(ref["value"].max(), 1e-30) + 1e-12
        # goal: equal when the oracle's maximum is clear of every cell that is not its immediate neighbour
        gr, gc = ref["global_goals"][0]
        v = ref["value"].copy()
        best = v[gr, gc]
        v[max(gr - 2, 0):gr + 3, max(gc - 2, 0):gc + 3] = -1
        got = tuple(int(x) for x in goals[e].cpu().tolist())
        if best - v.max() > 1e-2 * best:
            assert abs(got[0] - gr) <= 2 and abs(got[1] - gc) <= 2, (got, (gr, gc))
        assert val[got] >= best * (1 - 1e-2)
        assert int(st.goal_kind[e]) == 2 and int(st.last_kind[e]) == 1 and st.last_global_goal[e].cpu().tolist() == [3, 4]
        assert int(st.dd_wt_valid[e]) == 1


def test_goal_bookkeeping_and_stuck_rule(ctx):
    n, E = 240, 2
    st, host, dev = make_state(ctx, E, n, seeds=[21, 22])
    # environment 1: the agent is walled in (a solid block around it, nothing visited) -> only its own cell is reachable
    host["m"][1][:] = 0
    ar, ac = host["loc"][1][0] + host["lmb"][1][0], host["loc"][1][1] + host["lmb"][1][2]
    host["m"][1][max(ar - 12, 0):ar + 13, max(ac - 12, 0):ac + 13] = 1.0
    host["vis"][1][:] = 0
    host["coll"][1][:] = 0
    st.full_map[1, 0] = torch.from_numpy(host["m"][1])
    dev["vis"][1].zero_()
    dev["coll"][1].zero_()
    prev = torch.full((n // 2, n // 2), 0.25, dtype=torch.float64, device="cuda")
    st.update_global_goal(dev["tp"], 500.0, dev["coll"], dev["vis"], only_distance=True)   # allocates the state tensors
    st.dd_wt[1].copy_(prev)
    st.dd_wt_valid[1] = 1
    g1 = st.update_global_goal(dev["tp"], 500.0, dev["coll"], dev["vis"]).clone()
    torch.cuda.synchronize()
    assert int(torch.isfinite(st.dd[1]).sum()) == 1
    assert torch.equal(st.dd_wt[1], prev)                                # "stuck inside obstacle, use last dd_wt"
    ref1 = oracle_env(host, 1, prev_dd_wt=prev.cpu().numpy(), global_goals=[[3, 4]], last_global_goal=None)
    assert tuple(g1[1].cpu().tolist()) == tuple(int(x) for x in ref1["global_goals"][0])
    # second call, same inputs: the argmax equals the CURRENT goal (kind 2) but the LAST goal is the initial list (kind 1):
    # taken again, last <- current
    g2 = st.update_global_goal(dev["tp"], 500.0, dev["coll"], dev["vis"]).clone()
    assert torch.equal(g2, g1) and torch.equal(st.last_global_goal, g1) and st.last_kind.cpu().tolist() == [2, 2]
    # third call: equals the last goal (kind 2) -> suppressed; put a marker goal in to see that it is left alone
    st.global_goals.copy_(torch.tensor([[7, 7], [8, 8]], dtype=torch.int32))
    st.goal_kind.fill_(1)
    g3 = st.update_global_goal(dev["tp"], 500.0, dev["coll"], dev["vis"]).clone()
    torch.cuda.synchronize()
    assert g3.cpu().tolist() == [[7, 7], [8, 8]] and st.goal_kind.cpu().tolist() == [1, 1]
    assert torch.equal(st.last_global_goal, g1)


@pytest.mark.parametrize("temp", [-1.0, 0.0])
def test_special_temperatures(ctx, temp):
    st, host, dev = make_state(ctx, 1, 240, seeds=[31])
    g = st.update_global_goal(dev["tp"], temp, dev["coll"], dev["vis"])
    torch.cuda.synchronize()
    with np.errstate(divide="ignore"):
        ref = oracle_env(host, 0, dist_weight_temperature=temp, global_goals=[[3, 4]])
    val = st.value[0].cpu().numpy()
    if temp == -1.0:
        assert np.array_equal(val, host["tp"][0].astype(np.float64))
        assert tuple(g[0].cpu().tolist()) == tuple(int(x) for x in ref["global_goals"][0])
    else:
        # frontier mode cuts at dd < 60 (agent_state.py:405): a cell whose distance is within the field tolerance of the cut
        # can fall on either side of it (value 0 vs exp(-0.6)), everything else follows the 5e-3 weight tolerance
        lmb = host["lmb"][0]
        rd = ref["dd"][lmb[0]:lmb[1], lmb[2]:lmb[3]]
        gd = st.dd[0].cpu().numpy()[lmb[0]:lmb[1], lmb[2]:lmb[3]]        # the oracle's copy has the cut applied already
        clear = ~(np.abs(rd - 60.0) <= 0.5) & ~(np.abs(gd - 60.0) <= 0.5)
        assert clear.mean() > 0.95
        assert np.abs(val - ref["value"])[clear].max() <= 5e-3
        assert val[tuple(g[0].cpu().tolist())] >= ref["value"][clear].max() * (1 - 1e-2)


def test_no_masked_cell_quirk(ctx):
    """With nothing masked the reference's fill value is unused and `dd[dd == dd.max()] = inf` hits the farthest reached
    cells themselves (agent_state.py:392-393); the weights must follow (exp(-inf) = 0 there)."""
    n = 240
    st = MapState(ctx, 1, map_size_cm=n * 5, map_resolution=5, global_downscaling=2)
    st.lmb.copy_(torch.tensor([[0, n // 2, 0, n // 2]], dtype=torch.int32))
    st.loc.copy_(torch.tensor([[60, 60]], dtype=torch.int32))
    tp = torch.ones((1, n // 2, n // 2), device="cuda")
    st.update_global_goal(tp, 500.0)
    torch.cuda.synchronize()
    z = np.zeros((n, n), np.float32)
    ref = G.update_global_goal(z, z, z, (0, n // 2, 0, n // 2), 60, 60, np.ones((n // 2, n // 2)), global_goals=[[0, 0]])
    assert int(np.isinf(ref["dd"]).sum()) >= 1
    wt = st.dd_wt[0].cpu().numpy()
    assert np.abs(wt - ref["dd_wt"]).max() <= 5e-3
    # raw field on the device keeps the finite value at the farthest cell; the post-processing is applied where it is used
    assert bool(torch.isfinite(st.dd).all())
