"""Generates tests/golden/semmap_*.npz by running the UNMODIFIED reference Semantic_Mapping
(/root/reference/nav/agent/mapping.py) on seeded synthetic inputs, next to oracle/mapper.py on the same
inputs, and refuses to write unless the two agree bit for bit.  Run in the build container only
(the reference tree is not available on the GPU box):  python tests/golden/make_semmap_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/nav")
# mapping.py imports matplotlib.pyplot without using it
mpl = types.ModuleType("matplotlib")
mpl.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules.setdefault("matplotlib", mpl)
sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)

from agent.mapping import Semantic_Mapping  # noqa: E402  (the reference)
from oracle import mapper as oracle  # noqa: E402

CASES = [  # (name, seed, scene, sem_density, argparse overrides)
    ("room0", 0, "room", 0.1, {}),
    ("room1", 1, "room", 0.3, {}),
    ("stairs", 2, "stairs", 0.1, {}),
    ("wall", 3, "wall", 0.2, {}),
    ("empty", 4, "empty", 0.0, {}),
    # (camera_height * 100 + 1) / 5 is an integer in Python double arithmetic only for some heights: 0.89 gives max_z = 26
    # (a float32 round-trip of the height gives 25), so this case pins the obstacle height band's upper edge
    ("room_cam089", 5, "room", 0.2, {"camera_height": 0.89}),
    ("room_hfov90", 6, "room", 0.2, {"hfov": 90.0, "camera_height": 0.84}),
    # every pixel at the minimum depth: voxel columns with more than 2 048 entries (the key-only column kernel), coordinates
    # exactly on cell boundaries (zero-weight corners)
    ("wall_near", 7, "wall_near", 0.2, {}),
    # 15 semantic categories: 16 features per voxel (the 32-lane voxel slots), 17 ego channels, mapping.py:106-108's <= 16 branch
    ("room_cat15", 8, "room", 0.2, {"num_sem_categories": 15}),
]


def main():
    torch.set_num_threads(1)
    for name, seed, scene, dens, over in CASES:
        args = oracle.default_args(**over)
        ref_module = Semantic_Mapping(args).eval()
        obs = oracle.synth_obs(seed, args, scene, dens)
        delta, maps, poses = oracle.synth_state(seed, args)
        # reference (B = 1)
        p_ref = torch.from_numpy(poses.copy())
        with torch.no_grad():
            fp_r, map_r, pose_r, cur_r = ref_module(torch.from_numpy(obs)[None], torch.from_numpy(delta),
                                                    torch.from_numpy(maps), p_ref, None)
        # oracle restatement
        p_or = torch.from_numpy(poses.copy())[None]
        fp_o, map_o, pose_o, cur_o = oracle.forward(torch.from_numpy(obs)[None], torch.from_numpy(delta)[None],
                                                    torch.from_numpy(maps)[None], p_or, args)
        assert fp_r.shape == (1, 100, 100) and torch.equal(fp_r[0], fp_o[0]), name + ": fp_map_pred differs"
        assert torch.equal(map_r, map_o[0]), name + ": map_pred differs"
        assert torch.equal(cur_r, cur_o[0]) and torch.equal(p_ref, p_or[0]), name + ": pose differs"
        assert torch.equal(pose_r, p_ref), name + ": pose_pred must alias the mutated poses_last"
        out = os.path.join(HERE, f"semmap_{name}.npz")
        # inputs are regenerated from the seed by the tests; only the reference outputs are stored
        np.savez_compressed(out, seed=seed, scene=scene, sem_density=dens, overrides=repr(over),
                            fp_map_pred=fp_r[0].numpy().astype(np.float16),  # values are exactly 0 or 1
                            map_pred_q=np.round(map_r.numpy() * 65535.0).astype(np.uint16),  # 16-bit quantised copy
                            map_pred_sum=np.float64(map_r.double().sum().item()),
                            map_nonzero=np.int64((map_r != 0).sum().item()),
                            pose=cur_r.numpy())
        print(name, "ok", "fp cells", int(fp_r.sum()), "map nonzero", int((map_r != 0).sum()), os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
