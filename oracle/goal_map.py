"""Oracle for §8(f) N1c: ``Agent_State.update_goal_map`` (nav/agent/agent_state.py:423-452) - runs every step.

TEST INFRASTRUCTURE ONLY.  Pinned: tests/golden/make_goal_map_golden.py cuts the UNMODIFIED method out of the reference file
with ``ast`` and executes it on a stub state; bit-equality of goal_map (values AND dtype) and found_goal with this
restatement is required before tests/golden/goal_map.npz is written.  scikit-image is absent from the container: the two
functions the method calls are restated in the generator from scikit-image's published source (0.18 - 0.22 agree),

    skimage.morphology.binary_erosion(image)  == scipy.ndimage.binary_erosion(image, cross, border_value=True)
    skimage.morphology.binary_dilation(image) == scipy.ndimage.binary_dilation(image, cross)
    cross = scipy.ndimage.generate_binary_structure(2, 1)        (the default footprint)

and scipy.ndimage itself (present here) executes them, so only those two wrappers are unpinned.

Restated semantics:
  :429-430  goal_map = zeros (float64) with a 1 at global_goals[0]; found_goal = 0
  :433-436  if only_explore == 0 and local_map[goal_cat + 4] has a non-zero sum:
  :437-440      temp_goal = local_map[cn] with every value > 0 set to 1 (float32)
  :443-446      unless 'tv' is in the goal name: goal_erode x binary_erosion, then one binary_dilation (-> float64)
  :448          temp_goal *= (sum(local_map[4:10]) - local_map[cn]) == 0      (no OTHER category on the cell)
  :450-452      if temp_goal has a non-zero sum: goal_map = temp_goal, found_goal = 1
"""
import numpy as np
from scipy import ndimage as ndi

CROSS = ndi.generate_binary_structure(2, 1)


def update_goal_map(local_map, goal_cat, global_goal, goal_name, goal_erode=3, only_explore=0):
    """local_map [C, w, h] float32; returns (goal_map [w, h], found_goal)."""
    w, h = local_map.shape[1], local_map.shape[2]
    found_goal = 0
    goal_map = np.zeros((w, h))
    goal_map[global_goal[0], global_goal[1]] = 1
    if only_explore == 0:
        cn = goal_cat + 4
        if local_map[cn].sum(dtype=np.float32) != 0.:
            temp_goal = local_map[cn].copy()
            temp_goal[temp_goal > 0] = 1.
            if "tv" not in goal_name:
                for _ in range(goal_erode):
                    temp_goal = ndi.binary_erosion(temp_goal.astype(bool), structure=CROSS, border_value=True).astype(float)
                temp_goal = ndi.binary_dilation(temp_goal.astype(bool), structure=CROSS).astype(float)
            others = local_map[4].copy()
            for c in range(5, 10):  # torch.sum(local_map[4:10], dim=0): channel after channel, float32
                others = others + local_map[c]
            temp_goal *= (others - local_map[cn]) == 0
            if temp_goal.sum() != 0.:
                goal_map = temp_goal
                found_goal = 1
    return goal_map, found_goal


def synth_local_map(seed, nc=14, n=96, blobs=6, goal_cat=1):
    """Local map with a few rectangular / speckled category blobs; the goal category gets a thick blob (survives the
    erosion), a thin one (does not), one touching the border (border_value=True keeps it) and one overlapped by another
    category (masked out)."""
    rng = np.random.default_rng(seed)
    m = np.zeros((nc, n, n), np.float32)
    m[0] = (rng.random((n, n)) < 0.05)
    m[1] = (rng.random((n, n)) < 0.5)
    cn = goal_cat + 4
    for _ in range(blobs):
        c = int(rng.integers(4, nc))
        r0, c0 = int(rng.integers(0, n - 12)), int(rng.integers(0, n - 12))
        hh, ww = int(rng.integers(2, 14)), int(rng.integers(2, 14))
        m[c, r0:r0 + hh, c0:c0 + ww] = rng.random((min(hh, n - r0), min(ww, n - c0))).astype(np.float32) * 0.9 + 0.1
    m[cn, 10:22, 30:44] = 0.7                       # thick
    m[cn, 40:42, 10:30] = 0.5                       # thin: eroded away
    m[cn, 0:9, 60:75] = 1.0                         # touches the top border
    m[cn, n - 10:n, n - 11:n] = 0.3                 # corner
    m[cn, 60:75, 60:75] = 0.9
    other = 4 + (goal_cat + 1) % 6
    m[other, 64:70, 58:80] = 0.4                    # another category across the last blob
    m[cn] *= (rng.random((n, n)) > 0.01)            # pinholes
    return m
