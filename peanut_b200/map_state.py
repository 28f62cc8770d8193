"""Device-resident map bookkeeping of ``Agent_State`` (nav/agent/agent_state.py) - SURVEY.md section 8(f), N3.

Mirrors, for E environments at once, the fields and methods of the reference class that sit between the mapper (stage B)
and the planner: ``full_map`` / ``local_map``, ``full_pose`` / ``local_pose``, ``origins``, ``lmb``,
``planner_pose_inputs``, ``loc_r`` / ``loc_c``, ``dist_to_goal`` and

    init_map_and_pose()      agent_state.py:180-211
    stamp_initial()          agent_state.py:116-122   (the stamp of init_with_obs, after the first mapper call)
    update_local_map()       agent_state.py:276-303   (everything after the sem_map_module call)
    update_full_map()        agent_state.py:308-338
    update_global_goal()     agent_state.py:376-416   (obstacle dilation, geodesic distance field, exp weighting, argmax goal)
    update_goal_map()        agent_state.py:423-452   (goal found on the map? erode / dilate / no-other-category mask)

The reference computes every cell index on the host from ``pose.cpu().numpy()`` (three blocking device syncs per step and
environment) and applies the stamps as small indexed tensor writes.  Here every field is a CUDA tensor, each method is one
call into libpeanut_b200.so (``pn_map_*``) for all environments, and nothing synchronises; ``planner_inputs()`` is the one
place that reads the small state back.  There is no CPU path.
"""
import ctypes

import torch

from . import _lib


class MapState:
    """E environments' worth of Agent_State map fields on one device.  ``ctx`` is a ``_lib.Context``."""

    def __init__(self, ctx, num_envs, num_sem_categories=10, map_size_cm=4800, map_resolution=5, global_downscaling=2,
                 grid_resolution=24, col_rad=4, goal_reached_dist=75.0, f64_cells=False, device=None):
        if float(col_rad) != int(col_rad):
            raise ValueError("col_rad must be integer-valued (the reference indexes with selem_idx - (col_rad + 1))")
        self.ctx = ctx
        self.E = int(num_envs)
        self.nc = 4 + int(num_sem_categories)                                  # agent_state.py:39
        self.full_w = self.full_h = int(map_size_cm) // int(map_resolution)    # :41-42
        self.local_w = int(self.full_w / global_downscaling)                   # :43-44
        self.local_h = int(self.full_h / global_downscaling)
        dev = torch.device(device if device is not None else f"cuda:{ctx.device}")
        E = self.E
        self.full_map = torch.zeros((E, self.nc, self.full_w, self.full_h), dtype=torch.float32, device=dev)
        self.local_map = torch.zeros((E, self.nc, self.local_w, self.local_h), dtype=torch.float32, device=dev)
        self.full_pose = torch.zeros((E, 3), dtype=torch.float32, device=dev)
        self.local_pose = torch.zeros((E, 3), dtype=torch.float32, device=dev)
        self.origins = torch.zeros((E, 3), dtype=torch.float64, device=dev)
        self.lmb = torch.zeros((E, 4), dtype=torch.int32, device=dev)
        self.planner_pose_inputs = torch.zeros((E, 7), dtype=torch.float64, device=dev)
        self.loc = torch.zeros((E, 2), dtype=torch.int32, device=dev)          # loc_r, loc_c
        self.dist_to_goal = torch.zeros((E,), dtype=torch.float64, device=dev)
        self.global_goals = torch.zeros((E, 2), dtype=torch.int32, device=dev)  # global_goals[0] per environment
        self.cfg = _lib.MapCfg(self.nc, self.full_w, self.full_h, self.local_w, self.local_h, int(map_resolution),
                               int(map_size_cm), int(global_downscaling), int(grid_resolution), int(col_rad),
                               float(goal_reached_dist), 1 if f64_cells else 0)

    def _arrays(self):
        for t in (self.full_map, self.local_map, self.full_pose, self.local_pose, self.origins, self.lmb,
                  self.planner_pose_inputs, self.loc, self.dist_to_goal, self.global_goals):
            if not (t.is_cuda and t.is_contiguous()):
                raise TypeError("MapState fields must stay contiguous CUDA tensors")
        if tuple(self.local_map.shape) != (self.E, self.nc, self.local_w, self.local_h) or self.local_map.dtype != torch.float32:
            raise TypeError("local_map must be float32 [E, nc, local_w, local_h]")
        return _lib.MapArrays(self.full_map.data_ptr(), self.local_map.data_ptr(), self.full_pose.data_ptr(),
                              self.local_pose.data_ptr(), self.origins.data_ptr(), self.lmb.data_ptr(),
                              self.planner_pose_inputs.data_ptr(), self.loc.data_ptr(), self.dist_to_goal.data_ptr(),
                              self.global_goals.data_ptr())

    def _call(self, fn):
        arrays = self._arrays()
        stream = torch.cuda.current_stream(self.full_map.device).cuda_stream
        _lib.check(fn(self.ctx.handle, ctypes.byref(self.cfg), ctypes.byref(arrays), self.E, ctypes.c_void_p(stream)))

    def init_map_and_pose(self):
        self._call(self.ctx.lib.pn_map_init)

    def stamp_initial(self):
        self._call(self.ctx.lib.pn_map_stamp_initial)

    def update_local_map(self, local_map=None, local_pose=None):
        """Bookkeeping after the mapper call.  ``local_map`` / ``local_pose``: the mapper's outputs when it did not write
        into ``self.local_map`` / ``self.local_pose`` directly (adopted without a copy if contiguous)."""
        if local_map is not None:
            self.local_map = local_map.contiguous()
        if local_pose is not None:
            self.local_pose = local_pose.contiguous()
        self._call(self.ctx.lib.pn_map_update_local)

    def update_full_map(self):
        self._call(self.ctx.lib.pn_map_update_full)

    def update_goal_map(self, goal_cat, goal_names, goal_erode=3, only_explore=0):
        """``goal_cat``: per-environment goal category ids (sequence of ints or an int32 CUDA tensor [E]); ``goal_names``:
        the ``infos['goal_name']`` strings (erosion / dilation are skipped where 'tv' is in the name) or an int32 CUDA tensor
        of 0/1 skip flags.  Sets and returns ``self.goal_map`` [E, local_w, local_h] float32 and ``self.found_goal`` [E]
        int32, both on the device (the reference's goal_map is float64, or float32 in the 'tv' branch: same 0/1 values)."""
        dev = self.full_map.device
        E = self.E
        if not hasattr(self, "goal_map"):
            self.goal_map = torch.empty((E, self.local_w, self.local_h), dtype=torch.float32, device=dev)
            self.found_goal = torch.zeros((E,), dtype=torch.int32, device=dev)
        if only_explore != 0:                                                  # agent_state.py:433
            self.found_goal.zero_()
            self.goal_map.zero_()
            idx = torch.arange(E, device=dev)
            self.goal_map[idx, self.global_goals[:, 0].long(), self.global_goals[:, 1].long()] = 1.0
            return self.goal_map, self.found_goal
        if not torch.is_tensor(goal_cat):
            goal_cat = torch.tensor([int(g) for g in goal_cat], dtype=torch.int32, device=dev)
        if not torch.is_tensor(goal_names):
            goal_names = torch.tensor([1 if "tv" in str(n) else 0 for n in goal_names], dtype=torch.int32, device=dev)
        for t in (goal_cat, goal_names):
            if not (t.is_cuda and t.dtype == torch.int32 and t.numel() == E and t.is_contiguous()):
                raise TypeError("goal_cat / goal_names must be int32 CUDA tensors of E elements (or host sequences)")
        self._arrays()  # validates the state tensors
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self.ctx.lib.pn_goal_map(self.ctx.handle, self.local_map.data_ptr(), E, self.nc, self.local_w, self.local_h,
                                            goal_cat.data_ptr(), goal_names.data_ptr(), self.global_goals.data_ptr(),
                                            int(goal_erode), self.goal_map.data_ptr(), self.found_goal.data_ptr(),
                                            ctypes.c_void_p(stream)))
        return self.goal_map, self.found_goal

    def update_global_goal(self, target_pred, dist_weight_temperature=500.0, collision_map=None, visited_vis=None,
                           only_distance=False):
        """``Agent_State.update_global_goal`` for all environments (N1b).  ``target_pred``: float32 CUDA tensor
        [E, local_w, local_h] (``agent_prediction.update_prediction(..., as_numpy=False)`` per environment);
        ``collision_map`` / ``visited_vis``: uint8 CUDA tensors [E, full_w, full_h] (1 = set) or None - the reference
        keeps them on the host in the planner helper (agent_helper.py), here the caller uploads them when they change.
        Updates ``self.global_goals`` (and ``goal_kind`` / ``last_global_goal`` / ``last_kind`` / ``dd_wt``) in place and sets
        ``self.dd`` [E, full_w, full_h] float64 (inf = not traversible / unreachable) and ``self.value`` [E, local_w, local_h]
        float64.  Uses ``self.loc`` as written by ``update_local_map``.  Returns ``self.global_goals``."""
        dev = self.full_map.device
        E = self.E
        if not hasattr(self, "dd"):
            self.dd = torch.empty((E, self.full_w, self.full_h), dtype=torch.float64, device=dev)
            self.dd_wt = torch.zeros((E, self.local_w, self.local_h), dtype=torch.float64, device=dev)
            self.dd_wt_valid = torch.zeros((E,), dtype=torch.int32, device=dev)          # agent_state.py:88 (None)
            self.value = torch.zeros((E, self.local_w, self.local_h), dtype=torch.float64, device=dev)
            self.goal_kind = torch.ones((E,), dtype=torch.int32, device=dev)             # :127-129 list of lists
            self.last_global_goal = torch.zeros((E, 2), dtype=torch.int32, device=dev)
            self.last_kind = torch.zeros((E,), dtype=torch.int32, device=dev)            # :89 (None)
        if not only_distance:
            if not (torch.is_tensor(target_pred) and target_pred.is_cuda and target_pred.dtype == torch.float32 and
                    tuple(target_pred.shape) == (E, self.local_w, self.local_h) and target_pred.is_contiguous()):
                raise TypeError("target_pred must be a contiguous float32 CUDA tensor [E, local_w, local_h]")
        for name, t in (("collision_map", collision_map), ("visited_vis", visited_vis)):
            if t is not None and not (t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous() and
                                      tuple(t.shape) == (E, self.full_w, self.full_h)):
                raise TypeError(f"{name} must be a contiguous uint8 CUDA tensor [E, full_w, full_h]")
        self._arrays()
        cfg = _lib.GoalCfg(self.nc, self.full_w, self.full_h, self.local_w, self.local_h, int(self.cfg.col_rad),
                           int(self.cfg.map_resolution), float(dist_weight_temperature))
        ptr = lambda t: None if t is None else t.data_ptr()
        arrays = _lib.GoalArrays(self.full_map.data_ptr(), ptr(collision_map), ptr(visited_vis), self.lmb.data_ptr(),
                                 self.loc.data_ptr(), None if only_distance else target_pred.data_ptr(), self.dd.data_ptr(),
                                 self.dd_wt.data_ptr(), self.dd_wt_valid.data_ptr(), self.value.data_ptr(),
                                 self.global_goals.data_ptr(), self.goal_kind.data_ptr(), self.last_global_goal.data_ptr(),
                                 self.last_kind.data_ptr())
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self.ctx.lib.pn_global_goal(self.ctx.handle, ctypes.byref(cfg), ctypes.byref(arrays), E,
                                               1 if only_distance else 0, ctypes.c_void_p(stream)))
        return self.global_goals

    def planner_inputs(self):
        """One readback of the small per-environment state (the reference's host-side fields)."""
        packed = torch.cat([self.planner_pose_inputs, self.loc.double(), self.dist_to_goal[:, None], self.lmb.double(),
                            self.origins], dim=1).cpu().numpy()
        return {"pose_pred": packed[:, :7], "loc": packed[:, 7:9].astype(int), "dist_to_goal": packed[:, 9],
                "lmb": packed[:, 10:14].astype(int), "origins": packed[:, 14:17]}
