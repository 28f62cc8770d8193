"""Oracle for the glue between stages A and B: ``Agent_Helper._preprocess_obs`` / ``_preprocess_depth``
(nav/agent/agent_helper.py:175-217), restated in numpy.  TEST INFRASTRUCTURE ONLY.

Pinned: tests/golden/make_preproc_golden.py extracts the UNMODIFIED ``_preprocess_depth`` source from the
reference file, executes it next to this restatement on seeded inputs and requires bit-equality before writing
tests/golden/preproc_depth.npz.
"""
import numpy as np


def preprocess_depth(depth, min_d, max_d):
    """agent_helper.py:197-217.  depth float32 [H, W, 1] in [0, 1] (0 = invalid) -> float32 [H, W] in cm."""
    depth = depth[:, :, 0] * 1
    invalid = depth == 0.
    mostly = invalid.mean(axis=0) > 0.9                      # per column (:200-206)
    col_max = depth.max(axis=0)
    fill = np.where(mostly, col_max, np.float32(100.0)).astype(depth.dtype)
    depth = np.where(invalid, fill[None, :], depth)
    depth[depth > 0.99] = 0.                                  # too far (:208-210)
    depth[depth == 0] = 100.0                                 # (:212-213)
    return min_d * 100.0 + depth * (max_d - min_d) * 100.0    # (:215-216)


def preprocess_obs(rgb, depth, sem_seg_pred, env_frame_width=640, frame_width=160, min_d=0.5, max_d=5.0):
    """agent_helper.py:175-195 after the segmentation call -> float32 [4 + S, h, w].
    rgb uint8 [H,W,3], depth float32 [H,W,1], sem_seg_pred float32 [H,W,S].  The reference resizes rgb with
    PIL NEAREST, which for an integer factor picks the same pixels as [ds//2::ds]."""
    depth = preprocess_depth(depth, min_d, max_d)
    ds = env_frame_width // frame_width
    if ds != 1:
        rgb = rgb[ds // 2::ds, ds // 2::ds]
        depth = depth[ds // 2::ds, ds // 2::ds]
        sem_seg_pred = sem_seg_pred[ds // 2::ds, ds // 2::ds]
    depth = np.expand_dims(depth, axis=2)
    return np.concatenate((rgb, depth, sem_seg_pred), axis=2).transpose(2, 0, 1).astype(np.float32)


def synth_depth(seed, H=480, W=640):
    """Habitat-style normalised depth: smooth ramp + noise, 2 % exact zeros, 3 % > 0.99,
    plus a band of (almost) fully invalid columns to exercise the column-max branch."""
    rng = np.random.default_rng(seed)
    v = np.linspace(0.9, 0.15, H, dtype=np.float32)[:, None]
    d = v + 0.1 * np.sin(np.arange(W, dtype=np.float32)[None, :] / W * 6.0) + rng.normal(0, 0.01, (H, W)).astype(np.float32)
    d = np.clip(d, 0.02, 0.985).astype(np.float32)
    d[rng.random((H, W)) < 0.02] = 0.0
    d[rng.random((H, W)) < 0.03] = np.float32(0.995)
    c0 = int(rng.integers(0, W - 40))
    d[:, c0:c0 + 24] = 0.0
    d[rng.integers(0, H, 30), c0 + rng.integers(0, 24, 30)] = rng.uniform(0.1, 1.0, 30).astype(np.float32)
    return d[:, :, None]
