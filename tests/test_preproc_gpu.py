"""GPU: pn_make_obs (Agent_Helper._preprocess_obs / _preprocess_depth on the device) must be BIT-EXACT against the
numpy oracle, which is pinned bit-exact to the reference function (tests/golden/preproc_depth.npz)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import preproc as P
from peanut_b200 import _lib

pytestmark = pytest.mark.gpu


def _make_obs(ctx, depth, rgb, sem):
    E, H, W = depth.shape
    obs = torch.empty((E, 4 + sem.shape[-1], H // 4, W // 4), dtype=torch.float32, device="cuda")
    _lib.check(ctx.lib.pn_make_obs(ctx.handle, depth.data_ptr(), None if rgb is None else rgb.data_ptr(), sem.data_ptr(),
                                   E, H, W, H // 4, W // 4, sem.shape[-1], 0.5, 5.0, obs.data_ptr(), None))
    torch.cuda.synchronize()
    return obs.cpu().numpy()


def test_make_obs_bit_exact(ctx):
    rng = np.random.default_rng(0)
    depth = np.stack([P.synth_depth(s) for s in (0, 1, 2, 7)])           # [E,480,640,1]
    depth[3, :, 100] = 0.0                                                # a fully invalid column: max() is 0 -> 100
    rgb = rng.integers(0, 256, (4, 480, 640, 3)).astype(np.uint8)
    sem = (rng.random((4, 480, 640, 10)) < 0.1).astype(np.float32) * rng.integers(1, 3, (4, 480, 640, 10)).astype(np.float32)
    got = _make_obs(ctx, torch.from_numpy(depth[..., 0]).cuda(), torch.from_numpy(rgb).cuda(), torch.from_numpy(sem).cuda())
    for e in range(4):
        ref = P.preprocess_obs(rgb[e], depth[e], sem[e])
        assert np.array_equal(got[e], ref), f"env {e}: {(got[e] != ref).sum()} cells differ"
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "preproc_depth.npz"))
    for i in range(3):
        assert np.array_equal(got[i, 3], z["depth_cm_sub"][i])           # straight against the reference's output
    no_rgb = _make_obs(ctx, torch.from_numpy(depth[..., 0]).cuda(), None, torch.from_numpy(sem).cuda())
    assert (no_rgb[:, :3] == 0).all() and np.array_equal(no_rgb[:, 3:], got[:, 3:])
