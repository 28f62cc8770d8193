"""Times the device-resident map bookkeeping (pn_map_*) at the reference geometry (14 x 960^2 full, 14 x 480^2 local).

    python tools/map_state_profile.py [E]

Prints microseconds per call (CUDA events, 20 calls after 3 warm-ups) and the HBM rate of the window copies.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from peanut_b200 import _lib
from peanut_b200.map_state import MapState


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    ctx = _lib.Context(0)
    d = MapState(ctx, E)
    d.init_map_and_pose()
    d.local_map.uniform_(0, 1)
    local_bytes = d.local_map.numel() * 4
    t_local = timed(d.update_local_map)
    t_full = timed(d.update_full_map)
    t_init = timed(d.init_map_and_pose)
    d.global_goals.fill_(100)
    cats = torch.arange(E, dtype=torch.int32, device="cuda") % 6
    skip = torch.zeros(E, dtype=torch.int32, device="cuda")
    d.local_map[:, 4:] *= (d.local_map[:, 4:] > 0.7)
    t_goal = timed(lambda: d.update_goal_map(cats, skip))
    print(f"# map bookkeeping, E={E}, full {tuple(d.full_map.shape)}, local {tuple(d.local_map.shape)}")
    print(f"update_local_map  {t_local:8.1f} us   (writes channel 2: {local_bytes / d.nc / 1e6:.2f} MB)")
    print(f"update_full_map   {t_full:8.1f} us   ({4 * local_bytes / 1e6:.1f} MB moved -> {4 * local_bytes / t_full / 1e3:.0f} GB/s)")
    print(f"update_goal_map   {t_goal:8.1f} us   (reads 7 channels: {7 * local_bytes / d.nc / 1e6:.1f} MB, writes {local_bytes / d.nc / 1e6:.2f} MB)")
    print(f"init_map_and_pose {t_init:8.1f} us   (memset {d.full_map.numel() * 4 / 1e6:.1f} MB + window)")


if __name__ == "__main__":
    main()
