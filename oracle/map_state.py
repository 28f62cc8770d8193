"""Oracle for §8(f) N3: the map bookkeeping of ``Agent_State`` (nav/agent/agent_state.py:153-211, 112-122, 268-338).

TEST INFRASTRUCTURE ONLY.  Pinned: tests/golden/make_map_state_golden.py cuts the UNMODIFIED sources of
``get_local_map_boundaries``, ``init_map_and_pose``, ``update_local_map`` and ``update_full_map`` out of the reference file
with ``ast`` (the module itself imports skimage / skfmm / habitat, absent here), drives them over scripted trajectories on a
stub object and requires bit-equality of every map, pose and scalar with this restatement after every step before it
writes tests/golden/map_state.npz.

Restated semantics:
  get_local_map_boundaries :153-177  window origin snapped down to a multiple of grid_resolution, clamped into the full map
  init_map_and_pose        :180-211  zero maps, pose = map centre, 3x3 stamp on full_map[2:4], first window, local pose
  init_with_obs (stamp)    :116-122  3x3 stamp on local_map[2:4] at the cell of the local pose
  update_local_map (tail)  :276-303  planner pose, reset channel 2, 5x5 trajectory stamp on [2:4], explored disk of radius
                                     col_rad+1 under the agent (and at the goal once dist_to_goal < goal_reached_dist)
  update_full_map          :308-338  write the window back, recentre the window on the agent, re-cut local map and pose

Index semantics kept from the reference: slices follow Python rules (a negative start counts from the end, so a stamp that
pokes over the low edge is EMPTY, one over the high edge is clipped); the disk is written through integer index arrays, where
negative indices wrap and an index >= size raises IndexError in the reference (skipped here and on the device).

Cell arithmetic ``int(r * 100.0 / map_resolution)``: ``r`` is a numpy float32 scalar.  Under numpy >= 2 (NEP 50, what runs
in this container) the product and quotient stay float32; numpy < 2 promoted them to float64.  ``f64_cells`` selects the
second behaviour; the goldens are generated with the first.
"""
import numpy as np


def disk_idx(radius):
    """np.where(skimage.morphology.disk(radius) > 0): cells with x^2 + y^2 <= radius^2 on the (2r+1)^2 grid."""
    ax = np.arange(-radius, radius + 1)
    xx, yy = np.meshgrid(ax, ax)
    return np.where((xx ** 2 + yy ** 2) <= radius ** 2)


def _window_start(centre, local_n, full_n, grid):
    """One axis of the window: centred on the agent, snapped down to the grid, pushed back inside the full map."""
    start = centre - local_n // 2
    start -= start % grid            # Python modulo: non-negative, so this snaps DOWN also for negative starts
    if start < 0:
        start = 0
    if start + local_n > full_n:
        start = full_n - local_n
    return start


def boundaries(loc_r, loc_c, local_w, local_h, full_w, full_h, global_downscaling, grid_resolution):
    """[row0, row1, col0, col1] of the local window (agent_state.py:153-177)."""
    if global_downscaling <= 1:
        return [0, full_w, 0, full_h]
    r0 = _window_start(loc_r, local_w, full_w, grid_resolution)
    c0 = _window_start(loc_c, local_h, full_h, grid_resolution)
    return [r0, r0 + local_w, c0, c0 + local_h]


class MapState:
    """numpy mirror of the fields of Agent_State that the bookkeeping touches."""

    def __init__(self, nc, map_size_cm, map_resolution, global_downscaling, grid_resolution, col_rad, goal_reached_dist,
                 f64_cells=False):
        self.nc = nc
        self.map_size_cm, self.map_resolution = map_size_cm, map_resolution
        self.global_downscaling, self.grid_resolution = global_downscaling, grid_resolution
        self.col_rad, self.goal_reached_dist = col_rad, goal_reached_dist
        self.f64_cells = f64_cells
        self.full_w = self.full_h = map_size_cm // map_resolution
        self.local_w = int(self.full_w / global_downscaling)
        self.local_h = int(self.full_h / global_downscaling)
        self.full_map = np.zeros((nc, self.full_w, self.full_h), np.float32)
        self.local_map = np.zeros((nc, self.local_w, self.local_h), np.float32)
        self.full_pose = np.zeros(3, np.float32)
        self.local_pose = np.zeros(3, np.float32)
        self.origins = np.zeros(3)
        self.lmb = [0, 0, 0, 0]
        self.planner_pose_inputs = np.zeros(7)
        self.selem_idx = disk_idx(col_rad + 1)
        self.global_goals = [[0, 0]]
        self.loc_r = self.loc_c = 0
        self.dist_to_goal = 0.0

    def cell(self, v):
        v = np.float64(v) if self.f64_cells else np.float32(v)
        res = np.float64(self.map_resolution) if self.f64_cells else np.float32(self.map_resolution)
        hundred = np.float64(100.0) if self.f64_cells else np.float32(100.0)
        return int(v * hundred / res)

    def _recut(self, loc_r, loc_c):
        self.lmb = boundaries(loc_r, loc_c, self.local_w, self.local_h, self.full_w, self.full_h, self.global_downscaling,
                              self.grid_resolution)
        self.planner_pose_inputs[3:] = self.lmb
        self.origins = np.array([self.lmb[2] * self.map_resolution / 100.0, self.lmb[0] * self.map_resolution / 100.0, 0.])
        self.local_map = self.full_map[:, self.lmb[0]:self.lmb[1], self.lmb[2]:self.lmb[3]].copy()
        self.local_pose = self.full_pose - self.origins.astype(np.float32)

    def init_map_and_pose(self):
        self.full_map[:] = 0.
        self.full_pose[:] = 0.
        self.full_pose[:2] = self.map_size_cm / 100.0 / 2.0
        locs = self.full_pose
        self.planner_pose_inputs[:3] = locs
        loc_r, loc_c = self.cell(locs[1]), self.cell(locs[0])
        self.full_map[2:4, loc_r - 1:loc_r + 2, loc_c - 1:loc_c + 2] = 1.0
        self._recut(loc_r, loc_c)

    def stamp_initial(self):
        """init_with_obs, after the first mapper call (:116-122)."""
        loc_r, loc_c = self.cell(self.local_pose[1]), self.cell(self.local_pose[0])
        self.local_map[2:4, loc_r - 1:loc_r + 2, loc_c - 1:loc_c + 2] = 1.

    def _fill_disk(self, r0, c0):
        rr = self.selem_idx[0] - (self.col_rad + 1) + r0
        cc = self.selem_idx[1] - (self.col_rad + 1) + c0
        rr = np.asarray(rr).astype(np.int64)
        cc = np.asarray(cc).astype(np.int64)
        rr = np.where(rr < 0, rr + self.local_w, rr)
        cc = np.where(cc < 0, cc + self.local_h, cc)
        ok = (rr >= 0) & (rr < self.local_w) & (cc >= 0) & (cc < self.local_h)  # the reference raises IndexError otherwise
        self.local_map[1][rr[ok], cc[ok]] = 1.

    def update_local_map(self, new_local_map, new_local_pose):
        """The part of update_local_map after the mapper call; new_* are the mapper's outputs."""
        self.local_map = np.array(new_local_map, np.float32)
        self.local_pose = np.array(new_local_pose, np.float32)
        locs = self.local_pose
        self.planner_pose_inputs[:3] = locs + self.origins
        self.local_map[2, :, :] = 0.
        loc_r, loc_c = self.cell(locs[1]), self.cell(locs[0])
        traj_rad = 2
        self.local_map[2:4, loc_r - traj_rad:loc_r + traj_rad + 1, loc_c - traj_rad:loc_c + traj_rad + 1] = 1.
        self._fill_disk(loc_r, loc_c)
        g = self.global_goals[0]
        self.dist_to_goal = np.sqrt((loc_r - g[0]) ** 2 + (loc_c - g[1]) ** 2) * self.map_resolution
        if self.dist_to_goal < self.goal_reached_dist:
            self._fill_disk(g[0], g[1])
        self.loc_r, self.loc_c = loc_r, loc_c

    def update_full_map(self):
        self.full_map[:, self.lmb[0]:self.lmb[1], self.lmb[2]:self.lmb[3]] = self.local_map
        self.full_pose = self.local_pose + self.origins.astype(np.float32)
        locs = self.full_pose
        self._recut(self.cell(locs[1]), self.cell(locs[0]))
        locs = self.local_pose
        self.loc_r, self.loc_c = self.cell(locs[1]), self.cell(locs[0])


def scripted_mapper(rng, state, step_cells=3.0):
    """Deterministic stand-in for sem_map_module: sparse new observations max-ed into the map, pose advanced by a random
    step of up to `step_cells` cells.  Returns (local_map, local_pose) as float32 arrays."""
    lm = state.local_map.copy()
    seen = (rng.random(lm.shape) < 0.002).astype(np.float32) * rng.random(lm.shape).astype(np.float32)
    lm = np.maximum(lm, seen)
    step = (rng.random(2) * 2.0 - 1.0) * step_cells * state.map_resolution / 100.0
    pose = state.local_pose.copy()
    pose[:2] += step.astype(np.float32)
    pose[2] = np.float32(rng.random() * 360.0 - 180.0)
    return lm.astype(np.float32), pose.astype(np.float32)


# (name, nc, map_size_cm, map_resolution, global_downscaling, grid_resolution, col_rad, start offset (cells), steps,
#  step size (cells), num_local_steps, seed)
CASES = [
    ("centre", 6, 960, 5, 2, 24, 4, (0, 0), 24, 3.0, 5, 1),
    ("drift_far", 5, 960, 5, 2, 24, 4, (0, 0), 40, 9.0, 4, 2),       # walks far enough for the window to recentre
    ("corner_low", 5, 720, 5, 2, 24, 4, (-66, -66), 16, 2.0, 4, 3),  # window clamped at the low edges; stamps poke over the edge
    ("corner_high", 5, 720, 5, 2, 24, 4, (60, 64), 16, 2.0, 4, 4),   # window clamped at the high edges
    ("no_downscale", 5, 480, 5, 1, 24, 3, (0, 0), 10, 3.0, 5, 5),    # global_downscaling 1: local == full
    ("float_rad", 5, 960, 5, 2, 24, 4.0, (0, 0), 6, 3.0, 5, 6),      # col_rad as argparse delivers it (type=float)
    ("odd_grid", 5, 700, 5, 2, 10, 2, (9, -13), 12, 6.0, 3, 7),      # window origin not a multiple of 4 (scalar copy path)
]


def trajectory(case, f64_cells=False):
    """Drives a MapState through one scripted episode.  Yields (event, payload, state) AFTER the oracle applied the event:
    ("init", None), ("shift", shift[3] added to local_pose, followed by update_full_map), ("local", (local_map, local_pose,
    goal)) for an update_local_map with the mapper outputs and the current global goal, ("full", None)."""
    name, nc, size_cm, res, gds, grid, col_rad, start, steps, step_cells, nls, seed = case
    o = MapState(nc, size_cm, res, gds, grid, col_rad, 75.0, f64_cells=f64_cells)
    o.init_map_and_pose()
    yield "init", None, o
    shift = np.array([start[1] * res / 100.0, start[0] * res / 100.0, 0.0], np.float32)
    o.local_pose = o.local_pose + shift
    o.update_full_map()
    yield "shift", shift, o
    rng = np.random.default_rng(seed)
    for t in range(steps):
        lm, pose = scripted_mapper(rng, o, step_cells)
        if t == steps // 2:  # put the goal next to the agent: the second disk (goal marked explored) must fire
            o.global_goals = [[min(max(o.cell(pose[1]) + 3, 0), o.local_w - 1), min(max(o.cell(pose[0]) - 2, 0), o.local_h - 1)]]
        o.update_local_map(lm, pose)
        yield "local", (lm, pose, list(o.global_goals[0])), o
        if t % nls == nls - 1:
            o.update_full_map()
            yield "full", None, o


def digest(o):
    """Position-weighted checksums (float64) of every map channel + the small state."""
    w = np.arange(1, o.full_map[0].size + 1, dtype=np.float64).reshape(o.full_map[0].shape)
    wl = np.arange(1, o.local_map[0].size + 1, dtype=np.float64).reshape(o.local_map[0].shape)
    full = [(o.full_map[c].astype(np.float64) * w).sum() for c in range(o.nc)]
    local = [(o.local_map[c].astype(np.float64) * wl).sum() for c in range(o.nc)]
    small = list(o.full_pose.astype(np.float64)) + list(o.local_pose.astype(np.float64)) + list(o.origins) + \
        [float(v) for v in o.lmb] + list(o.planner_pose_inputs) + [float(o.loc_r), float(o.loc_c), float(o.dist_to_goal)]
    return np.array(full + local + small, np.float64)


INIT_POSES = [(12.03, 11.98), (0.04, 7.0), (7.0, 0.02), (23.97, 23.99), (3.3, 23.96)]  # local poses (m) for init_with_obs


def init_with_obs_case(k):
    """State after init_map_and_pose + a scripted first mapper call at INIT_POSES[k], BEFORE the 3x3 stamp."""
    o = MapState(5, 960, 5, 2, 24, 4, 75.0)
    o.init_map_and_pose()
    rng = np.random.default_rng(100 + k)
    lm, _ = scripted_mapper(rng, o)
    o.local_map, o.local_pose = lm.copy(), np.array([INIT_POSES[k][0], INIT_POSES[k][1], 10.0], np.float32)
    return o
