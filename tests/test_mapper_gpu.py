"""Stage B parity on the GPU, through the reference-facing Semantic_Mapping shim and the C-ABI.

* Integer work - voxel counts -> thresholded ego map (obstacle / explored / category cells), fp_map_pred,
  the stair-mask decision - must be BIT-EXACT against the oracle (which is pinned bit-exact to the reference).
* Floating point - pose update (sinf/cosf vs the CPU's), and the two chained bilinear resamplings -
  tolerance 1e-4 absolute on map cells (values in [0,1], gradient <= 1 per cell, sampling coordinates
  carry ~1 ulp * 479 of error per resampling), 1e-5 on poses.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mapper as oracle
from peanut_b200.mapping import Semantic_Mapping
from tests.test_mapper_oracle_cpu import GOLDEN, check_against_golden, load_case

pytestmark = pytest.mark.gpu
MAP_TOL = 1e-4
POSE_TOL = 1e-5


def gpu_args():
    a = oracle.default_args()
    a.device = torch.device("cuda:0")
    return a


@pytest.fixture(scope="module")
def module1():
    return Semantic_Mapping(gpu_args()).to("cuda:0").eval()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_reference_golden_single_env(module1, path):
    z, args, obs, delta, maps, poses = load_case(path)
    p = torch.from_numpy(poses.copy()).cuda()
    if "overrides" in z.files and str(z["overrides"]) != "{}":  # non-default camera geometry: its own module
        args.device = torch.device("cuda:0")
        module1 = Semantic_Mapping(args).to("cuda:0").eval()
    fp, mp, pose_pred, cur = module1(torch.from_numpy(obs)[None].cuda(), torch.from_numpy(delta).cuda(),
                                     torch.from_numpy(maps).cuda(), p, None)
    torch.cuda.synchronize()
    assert fp.shape == (1, 100, 100) and mp.shape == (4 + args.num_sem_categories, 480, 480)
    assert pose_pred is p and cur is p  # aliasing contract
    check_against_golden(z, fp[0].cpu().numpy(), mp.cpu().numpy(), cur.cpu().numpy(), map_tol=MAP_TOL, pose_tol=POSE_TOL)


def _run_both(cases, sem_density=0.2):
    args = oracle.default_args()
    obs = torch.stack([torch.from_numpy(oracle.synth_obs(s, args, sc, sem_density)) for s, sc in cases])
    st = [oracle.synth_state(s, args) for s, _ in cases]
    delta = torch.stack([torch.from_numpy(x[0]) for x in st])
    maps = torch.stack([torch.from_numpy(x[1]) for x in st])
    poses = torch.stack([torch.from_numpy(x[2]) for x in st])
    g = oracle.Geometry(args)
    fp_o, ego_o = oracle.ego_map(obs.clone(), g)
    fp_ref, map_ref, _, cur_ref = oracle.forward(obs.clone(), delta, maps, poses.clone(), args)
    mod = Semantic_Mapping(gpu_args(), num_envs=len(cases))
    p = poses.clone().cuda()
    fp, mp, _ = mod.forward_batch(obs.cuda(), delta.cuda(), maps.cuda(), p)
    ego, flags = mod.read_ego()
    return dict(fp=fp.cpu(), map=mp.cpu(), pose=p.cpu(), ego=ego.cpu(), flags=flags.cpu(), fp_ref=fp_ref,
                map_ref=map_ref, pose_ref=cur_ref, ego_ref=ego_o, obs=obs, g=g)


def test_batched_envs_exact_integer_part():
    cases = [(10, "room"), (11, "stairs"), (12, "wall"), (13, "empty"), (14, "room"), (15, "room"), (16, "stairs"), (17, "wall")]
    r = _run_both(cases)
    g = r["g"]
    # ego window of the oracle's agent_view: rows 240..339, cols 190..289; channels 0,1,4..13
    n = g.map_cells
    win = r["ego_ref"][:, :, n // 2:n // 2 + g.vr, n // 2 - g.vr // 2:n // 2 + g.vr // 2]
    ego_ref = torch.cat([win[:, 0:2], win[:, 4:]], 1)
    mism = (r["ego"] != ego_ref)
    assert not mism.any(), f"{int(mism.sum())} ego cells differ (of {mism.numel()})"
    assert torch.equal(r["fp"], r["fp_ref"])
    # stair flags: which envs had their low points masked
    expect = []
    for i in range(len(cases)):
        c = oracle.normalised_coords(r["obs"][i:i + 1, 3], g)
        before = c.clone()
        feat = torch.ones(1, g.nf, g.h * g.w)
        feat[:, 1:] = r["obs"][i:i + 1, 4:].reshape(1, -1, g.h * g.w)
        oracle.stair_mask_(c, feat)
        expect.append(int(not torch.equal(before, c)))
    assert r["flags"].tolist() == expect and 0 < sum(expect) < len(expect)
    assert (r["map"] - r["map_ref"]).abs().max().item() <= MAP_TOL
    assert (r["pose"] - r["pose_ref"]).abs().max().item() <= POSE_TOL
    # derived integer category map (compress_sem_map of the semantic channels, segmentation.py:65-69)
    def cat_map(m):
        out = torch.zeros(m.shape[0], m.shape[2], m.shape[3], dtype=torch.int64)
        for i in range(4, m.shape[1]):
            out[m[:, i] > 0.] = i - 3
        return out
    a, b = cat_map(r["map"]), cat_map(r["map_ref"])
    # cells whose value is within the float tolerance of zero may flip; everything else must agree exactly
    sure = (r["map_ref"][:, 4:].abs() > 2 * MAP_TOL).any(1) | (r["map_ref"][:, 4:] == 0).all(1)
    assert torch.equal(a[sure], b[sure])


def test_strided_maps_last_view_and_fresh_output():
    """maps_last as a window of a larger full_map (agent_state.py:206-208); map_pred must be a fresh tensor."""
    args = oracle.default_args()
    obs = torch.from_numpy(oracle.synth_obs(21, args, "room"))[None]
    delta, maps, poses = oracle.synth_state(21, args)
    full = torch.zeros((14, 960, 960))
    full[:, 100:580, 200:680] = torch.from_numpy(maps)
    full_d = full.cuda()
    view = full_d[:, 100:580, 200:680]
    assert not view.is_contiguous()
    mod = Semantic_Mapping(gpu_args())
    p = torch.from_numpy(poses.copy()).cuda()
    fp, mp, _, _ = mod(obs.cuda(), torch.from_numpy(delta).cuda(), view, p, None)
    p2 = torch.from_numpy(poses.copy())[None]
    _, map_ref, _, _ = oracle.forward(obs, torch.from_numpy(delta)[None], torch.from_numpy(maps)[None], p2, args)
    assert (mp.cpu() - map_ref[0]).abs().max().item() <= MAP_TOL
    assert mp.data_ptr() != view.data_ptr() and mp.is_contiguous()
    assert torch.equal(full_d.cpu(), full)  # input untouched


def test_idempotent_max_fuse():
    """Size-independent property: feeding map_pred back with zero motion and the same frame changes nothing
    (max-fusion is idempotent once the pose is fixed)."""
    args = oracle.default_args()
    obs = torch.from_numpy(oracle.synth_obs(31, args, "room"))[None].cuda()
    _, maps, poses = oracle.synth_state(31, args)
    mod = Semantic_Mapping(gpu_args())
    p = torch.from_numpy(poses.copy()).cuda()
    zero = torch.zeros(3, device="cuda")
    _, m1, _, _ = mod(obs, zero, torch.from_numpy(maps).cuda(), p, None)
    _, m2, _, _ = mod(obs, zero, m1, p, None)
    assert torch.equal(m1, m2)
    assert (m1 >= torch.from_numpy(maps).cuda()).all()


def test_wrong_shapes_raise():
    mod = Semantic_Mapping(gpu_args())
    with pytest.raises(TypeError):
        mod(torch.zeros((1, 14, 120, 161), device="cuda"), torch.zeros(3, device="cuda"),
            torch.zeros((14, 480, 480), device="cuda"), torch.zeros(3, device="cuda"), None)
