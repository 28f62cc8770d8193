"""One perception step for E independent environments on one GPU: RGB-D frame + partial map -> (per-category masks,
updated local map + pose, predicted semantic map).

Two orderings of the same kernels:

* ``mode="dependent"`` (default) is the reference's own chain - A (Mask-RCNN) -> glue -> B (mapper) -> the head of
  ``Agent_State.update_prediction`` (agent_state.py:350-360: stamp the UPDATED local map into the full map, cut the
  prediction window) -> C (map completion) - so the predicted map reflects the current frame.  The reference moves the
  window device -> host -> device between B and C; here stamp and crop are two device kernels.
* ``mode="overlapped"`` runs C on a side stream on the partial map the CALLER supplies (i.e. last step's map), next to A
  and B: higher throughput, but the prediction is one frame stale.  Not the reference's ordering; kept as an option.

This is the batched composition of the three reference call sites that ``PEANUT_Agent.act`` walks every step
(SURVEY.md §3.1): ``SemanticPredMaskRCNN.get_prediction`` (nav/agent/agent_helper.py:220-225),
``Agent_Helper._preprocess_obs`` (agent_helper.py:175-195), ``Semantic_Mapping.forward``
(nav/agent/agent_state.py:273-274) and ``PEANUT_Prediction_Model.get_prediction`` (agent_state.py:355-361).  The
reference runs them for one environment with >= 3 host->device and >= 6 device->host synchronising copies per step;
here everything between the input copy and the result copy stays on the device, on one stream, for E
environments at once (the environments are independent: no cross-env state).
"""
import ctypes
import types

import torch

from . import _lib
from .mapping import Semantic_Mapping
from .prediction import Segmentor, _default_cfg
from .segmentation import MaskRCNN


def default_args(**kw):
    """The argparse defaults of nav/arguments.py that shape the path (lines 40-81, 104)."""
    a = dict(env_frame_width=640, env_frame_height=480, frame_width=160, frame_height=120, min_depth=0.5, max_depth=5.0,
             map_resolution=5, map_size_cm=4800, global_downscaling=2, vision_range=100, hfov=79.0, du_scale=1,
             cat_pred_threshold=5.0, exp_pred_threshold=1.0, map_pred_threshold=0.1, num_sem_categories=10,
             camera_height=0.88, sem_pred_prob_thr=0.95, goal_thr=0.985, sem_gpu_id=0)
    a.update(kw)
    return types.SimpleNamespace(**a)


class PerceptionPipeline:
    def __init__(self, seg_weights, pred_weights, num_envs=1, device="cuda:0", precision="bf16", map_shape=(24, 240, 240),
                 num_pred_classes=6, args=None, mode="dependent"):
        if mode not in ("dependent", "overlapped"):
            raise ValueError("mode must be 'dependent' (reference ordering) or 'overlapped' (stale-map throughput mode)")
        self.mode = mode
        self.args = args or default_args()
        a = self.args
        dev = torch.device(device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else 0)
        a.device = self.device
        self.E = int(num_envs)
        self.map_shape = tuple(map_shape)
        self.num_pred_classes = num_pred_classes
        self.seg = MaskRCNN(seg_weights, device=self.device, precision=precision, batch=self.E, height=a.env_frame_height,
                            width=a.env_frame_width)
        self.mapper = Semantic_Mapping(a, num_envs=self.E)
        self._pred_weights, self._precision = pred_weights, precision
        self.pred = Segmentor(_default_cfg(map_shape[0], num_pred_classes), pred_weights, self.device, precision=precision)
        self.pred._ensure_built(self.E, *self.map_shape)
        nsem = a.num_sem_categories
        E, H, W = self.E, a.env_frame_height, a.env_frame_width
        d = self.device
        self.sem = torch.zeros((E, H, W, nsem), dtype=torch.float32, device=d)
        self.obs = torch.zeros((E, 4 + nsem, a.frame_height, a.frame_width), dtype=torch.float32, device=d)
        self.pred_out = torch.zeros((E, num_pred_classes) + self.map_shape[1:], dtype=torch.float32, device=d)
        # dependent mode: the full map the updated local map is stamped into, the (fixed, centred) local-map bounds and the
        # prediction window cut out of it (agent_state.py:186-204, 350-360)
        self.nc = 4 + nsem
        self.full_w = self.full_h = a.map_size_cm // a.map_resolution
        self.local_w = self.local_h = self.full_w // a.global_downscaling
        Cm, Hm, Wm = self.map_shape
        if mode == "dependent":
            if Hm > self.full_w or Wm > self.full_h:
                raise ValueError("the prediction window does not fit the full map")
            self.full_map = torch.zeros((E, self.nc, self.full_w, self.full_h), dtype=torch.float32, device=d)
            r0, c0 = (self.full_w - self.local_w) // 2, (self.full_h - self.local_h) // 2
            self.lmb = torch.tensor([[r0, r0 + self.local_w, c0, c0 + self.local_h]] * E, dtype=torch.int32, device=d)
            self.win_x1, self.win_y1 = self.full_w // 2 - Hm // 2, self.full_h // 2 - Wm // 2   # agent_state.py:357-360
            # channels of the net's input that exist in the map come from the window; a net with MORE input planes than
            # the map has (BASELINE's synthetic 24 x 240 x 240 against the reference's 14 map channels) keeps the caller's
            # planes there
            self.win_channels = min(Cm, self.nc)
        self._side = torch.cuda.Stream(device=d)
        self._fork, self._join, self._join_d2h = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self._depth_ev, self._h2d_fence, self._depth_ready = torch.cuda.Event(), torch.cuda.Event(), None
        self._pmap_ev, self._pmap_ready = torch.cuda.Event(), None
        # staging for the host entry point
        self._dev_in = None
        self._host_out = None
        if self.seg.n_cats + 1 != nsem:
            raise ValueError("num_sem_categories must equal the segmentation classes + 1")

    def launches_per_step(self):
        """Kernel launches one step enqueues (graph nodes of the two networks + mapper + glue [+ stamp and crop])."""
        return self.seg.num_launches() + 1 + int(self.mapper.ctx.lib.pn_semmap_num_launches(self.mapper.ctx.handle)) + \
            self.pred.num_launches() + (2 if self.mode == "dependent" else 0)

    def pred_single(self):
        """A batch-1 engine of the same map-completion net (same weights, precision), built on first use: what the
        single-environment reference call ``PEANUT_Prediction_Model.get_prediction`` runs."""
        if getattr(self, "_pred1", None) is None:
            self._pred1 = Segmentor(_default_cfg(self.map_shape[0], self.num_pred_classes), self._pred_weights, self.device,
                                    precision=self._precision)
        return self._pred1

    def _glue_mapper_window(self, rgb, depth, pose_delta, local_map, poses, partial_map, map_out=None, fp_out=None):
        """Everything between stage A and stage C of the dependent chain, on the current stream: `_preprocess_obs` glue,
        the mapper, and update_prediction's stamp + prediction window (agent_state.py:350-360)."""
        a = self.args
        main = torch.cuda.current_stream(self.device)
        if self._depth_ready is not None:  # step_host copies depth / pose delta beside Mask-RCNN
            main.wait_event(self._depth_ready)
        stream = ctypes.c_void_p(main.cuda_stream)
        lib, h = self.seg.ctx.lib, self.seg.ctx.handle
        _lib.check(lib.pn_make_obs(h, depth.data_ptr(), rgb.data_ptr(), self.sem.data_ptr(), self.E, a.env_frame_height,
                                   a.env_frame_width, a.frame_height, a.frame_width, a.num_sem_categories, a.min_depth,
                                   a.max_depth, self.obs.data_ptr(), stream))
        fp, new_map, poses = self.mapper.forward_batch(self.obs, pose_delta, local_map, poses, out=map_out, fp_out=fp_out)
        Cm, Hm, Wm = self.map_shape
        _lib.check(lib.pn_map_stamp_local(h, new_map.data_ptr(), self.full_map.data_ptr(), self.lmb.data_ptr(), self.E, self.nc,
                                          self.local_w, self.local_h, self.full_w, self.full_h, stream))
        if self._pmap_ready is not None:  # step_host copies the caller's partial map beside Mask-RCNN
            main.wait_event(self._pmap_ready)
        _lib.check(lib.pn_map_crop_window(h, self.full_map.data_ptr(), self.E, self.nc, self.full_w, self.full_h, self.win_x1,
                                          self.win_y1, Hm, Wm, self.win_channels, partial_map.data_ptr(), Cm, stream))
        return fp, new_map, poses

    def _step_dependent(self, rgb, depth, pose_delta, local_map, poses, partial_map, goal_cat):
        """A -> glue -> B -> stamp -> crop -> C on the caller's stream (the reference's chain, agent_state.py:273-274 then
        :350-361).  ``partial_map`` [E,C,Hm,Wm] is the net's input buffer: its first min(C, nc) planes are OVERWRITTEN with
        the prediction window of the updated full map."""
        a = self.args
        main = torch.cuda.current_stream(self.device)
        self.seg.forward_device(rgb, goal_cat, a.sem_pred_prob_thr, a.sem_pred_prob_thr, a.goal_thr, out=self.sem)
        fp, new_map, poses = self._glue_mapper_window(rgb, depth, pose_delta, local_map, poses, partial_map)
        pred = self.pred.forward_device(partial_map, apply_sigmoid=True, out=self.pred_out)
        return self.sem, fp, new_map, poses, pred

    def time_glue_mapper_window(self, rgb, depth, pose_delta, local_map, poses, partial_map, iters=5):
        """Device milliseconds of the kernels between stage A and stage C (CUDA events, after one warm-up): what the
        per-launch profiles of the two networks do not cover.  Dependent mode only."""
        if self.mode != "dependent":
            raise RuntimeError("time_glue_mapper_window: dependent mode only")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p = poses.clone()
        self._glue_mapper_window(rgb, depth, pose_delta, local_map, p, partial_map)
        e0.record()
        for _ in range(iters):
            self._glue_mapper_window(rgb, depth, pose_delta, local_map, p, partial_map)
        e1.record()
        torch.cuda.synchronize(self.device)
        return e0.elapsed_time(e1) / iters

    def step_device(self, rgb, depth, pose_delta, local_map, poses, partial_map, goal_cat=None):
        """All CUDA tensors: rgb uint8 [E,H,W,3]; depth float32 [E,H,W] (simulator units, 0 = invalid);
        pose_delta [E,3]; local_map [E,4+S,n,n]; poses [E,3] (updated in place); partial_map [E,C,Hm,Wm].
        Returns (sem [E,H,W,S], fp_map [E,vr,vr], new_local_map [E,4+S,n,n], poses, pred_map [E,K,Hm,Wm]);
        no host synchronisation."""
        if not (partial_map.is_cuda and partial_map.dtype == torch.float32 and partial_map.is_contiguous() and
                tuple(partial_map.shape) == (self.E,) + self.map_shape):
            raise TypeError(f"partial_map: expected a contiguous float32 CUDA tensor of shape {(self.E,) + self.map_shape}")
        if self.mode == "dependent":
            return self._step_dependent(rgb, depth, pose_delta, local_map, poses, partial_map, goal_cat)
        a = self.args
        # The map-completion net only reads the caller's partial map: it runs on a side stream next to Mask-RCNN and the
        # mapper (at small E every layer is a single wave of CTAs that leaves room for a second resident CTA per SM).
        main = torch.cuda.current_stream(self.device)
        self._fork.record(main)
        # the critical chain is enqueued first; the side stream's work follows a few microseconds of host time later
        self.seg.forward_device(rgb, goal_cat, a.sem_pred_prob_thr, a.sem_pred_prob_thr, a.goal_thr, out=self.sem)
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._fork)
            pred = self.pred.forward_device(partial_map, apply_sigmoid=True, out=self.pred_out)
            self._join.record(self._side)
        partial_map.record_stream(self._side)
        if self._depth_ready is not None:  # step_host copies depth / pose delta beside Mask-RCNN
            main.wait_event(self._depth_ready)
        stream = main.cuda_stream
        lib = self.seg.ctx.lib
        _lib.check(lib.pn_make_obs(self.seg.ctx.handle, depth.data_ptr(), rgb.data_ptr(), self.sem.data_ptr(), self.E,
                                   a.env_frame_height, a.env_frame_width, a.frame_height, a.frame_width,
                                   a.num_sem_categories, a.min_depth, a.max_depth, self.obs.data_ptr(),
                                   ctypes.c_void_p(stream)))
        fp, new_map, poses = self.mapper.forward_batch(self.obs, pose_delta, local_map, poses)
        main.wait_event(self._join)
        return self.sem, fp, new_map, poses, pred

    def step_host(self, rgb_h, depth_h, pose_delta_h, partial_map_h, local_map, poses):
        """End-to-end step as a caller with HOST observations sees it: pinned host inputs are copied to the device,
        the step runs, and the step's result (predicted map, pose, egocentric obstacle map) is copied back.
        The local map and poses are device-resident state, as in the reference (agent_state.py:53-60).
        Returns (pred_map_host, poses_host, fp_map_host, new_local_map_device)."""
        d = self.device
        if self._dev_in is None:
            self._dev_in = (torch.empty_like(rgb_h, device=d), torch.empty_like(depth_h, device=d),
                            torch.empty_like(pose_delta_h, device=d), torch.empty_like(partial_map_h, device=d))
            self._host_out = (torch.empty(self.pred_out.shape, dtype=torch.float32).pin_memory(),
                              torch.empty((self.E, 3), dtype=torch.float32).pin_memory(),
                              torch.empty((self.E, self.args.vision_range, self.args.vision_range), dtype=torch.float32).pin_memory())
        # Only the frame gates Mask-RCNN.  Depth and the pose delta (needed after it, by the glue and the mapper) and the
        # partial map (the largest input, read only by the map-completion net on the side stream) are copied on the side
        # stream, beside Mask-RCNN instead of ahead of it.
        self._dev_in[0].copy_(rgb_h, non_blocking=True)
        main = torch.cuda.current_stream(d)
        self._h2d_fence.record(main)   # earlier work on the caller's stream (e.g. the previous step) precedes the copies
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._h2d_fence)
            self._dev_in[1].copy_(depth_h, non_blocking=True)
            self._dev_in[2].copy_(pose_delta_h, non_blocking=True)
            self._depth_ev.record(self._side)
            self._dev_in[3].copy_(partial_map_h, non_blocking=True)
            self._pmap_ev.record(self._side)
        rgb, depth, delta, pmap = self._dev_in
        self._depth_ready = self._depth_ev
        self._pmap_ready = self._pmap_ev if self.mode == "dependent" else None
        try:
            _, fp, new_map, poses, pred = self.step_device(rgb, depth, delta, local_map, poses, pmap)
        finally:
            self._depth_ready = None
            self._pmap_ready = None
        if self.mode == "dependent":  # one chain on the caller's stream: the results leave in order behind stage C
            self._host_out[1].copy_(poses, non_blocking=True)
            self._host_out[2].copy_(fp, non_blocking=True)
            self._host_out[0].copy_(pred, non_blocking=True)
        else:
            with torch.cuda.stream(self._side):  # the predicted map leaves on the side stream as soon as it exists
                self._host_out[0].copy_(pred, non_blocking=True)
                self._join_d2h.record(self._side)
            self._host_out[1].copy_(poses, non_blocking=True)
            self._host_out[2].copy_(fp, non_blocking=True)
            main.wait_event(self._join_d2h)
        main.synchronize()
        return self._host_out[0], self._host_out[1], self._host_out[2], new_map

    def h2d_bytes(self, rgb_h, depth_h, pose_delta_h, partial_map_h):
        return sum(t.numel() * t.element_size() for t in (rgb_h, depth_h, pose_delta_h, partial_map_h))

    def d2h_bytes(self):
        return sum(t.numel() * t.element_size() for t in self._host_out) if self._host_out else 0


class MicroBatchedPipeline:
    """The dependent chain pipelined over micro-batches of environments on one GPU.

    Each environment's chain stays the reference's (Mask-RCNN -> glue -> mapper -> stamp + window -> map completion), but the
    E environments are split into ``micro_batches`` groups with their own engines: while the caller's stream runs Mask-RCNN of
    group i + 1, a side stream runs everything after Mask-RCNN of group i.  Mask-RCNN's persistent tensor-core kernels leave
    registers and warp slots idle that the latency-bound mapper kernels (low occupancy, little shared memory) can use, so
    that work leaves the critical path.  Results are written into slices of shared [E, ...] tensors (no concatenation, no
    allocation inside the step); the local map is double-buffered because the mapper's output must not alias its input.
    Same public surface as ``PerceptionPipeline`` (``step_device`` / ``step_host`` / byte and launch counts)."""

    def __init__(self, seg_weights, pred_weights, num_envs=2, micro_batches=2, device="cuda:0", precision="bf16",
                 map_shape=(24, 240, 240), num_pred_classes=6, args=None):
        if micro_batches < 2 or num_envs % micro_batches != 0:
            raise ValueError("num_envs must be a multiple of micro_batches >= 2")
        self.E, self.mb, self.Es = int(num_envs), int(micro_batches), int(num_envs) // int(micro_batches)
        self.mode = "dependent"
        self.subs = [PerceptionPipeline(seg_weights, pred_weights, self.Es, device, precision, map_shape, num_pred_classes,
                                        args=None if args is None else types.SimpleNamespace(**vars(args)), mode="dependent")
                     for _ in range(self.mb)]
        p0 = self.subs[0]
        self.device, self.args, self.map_shape = p0.device, p0.args, p0.map_shape
        d = self.device
        self.seg, self.pred = p0.seg, p0.pred       # (for callers that only need the shapes / one engine's profile)
        self.pred_out = torch.zeros((self.E,) + tuple(p0.pred_out.shape[1:]), dtype=torch.float32, device=d)
        self.fp_out = torch.zeros((self.E, p0.args.vision_range, p0.args.vision_range), dtype=torch.float32, device=d)
        shape = (self.E, p0.nc, p0.local_w, p0.local_h)
        self._maps = [torch.zeros(shape, dtype=torch.float32, device=d) for _ in range(2)]
        for i, s in enumerate(self.subs):
            s.pred_out = self.pred_out[i * self.Es:(i + 1) * self.Es]
        self._side = [torch.cuda.Stream(device=d) for _ in range(self.mb - 1)]
        self._a_done = [torch.cuda.Event() for _ in range(self.mb)]
        self._rest_done = [torch.cuda.Event() for _ in range(self.mb - 1)]
        self._copy = torch.cuda.Stream(device=d)
        self._fence, self._depth_ev, self._pmap_ev = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self._dev_in = None
        self._host_out = None

    def launches_per_step(self):
        return sum(s.launches_per_step() for s in self.subs)

    def step_device(self, rgb, depth, pose_delta, local_map, poses, partial_map, goal_cat=None, _events=(None, None)):
        """Same contract as PerceptionPipeline.step_device (dependent mode); ``sem`` is returned as a list of the groups'
        [Es,H,W,S] tensors, everything else as [E, ...] tensors."""
        a = self.args
        main = torch.cuda.current_stream(self.device)
        out_map = self._maps[0] if local_map.data_ptr() != self._maps[0].data_ptr() else self._maps[1]
        sems = []
        for i, sub in enumerate(self.subs):
            sl = slice(i * self.Es, (i + 1) * self.Es)
            g = None if goal_cat is None else goal_cat[sl]
            sub.seg.forward_device(rgb[sl], g, a.sem_pred_prob_thr, a.sem_pred_prob_thr, a.goal_thr, out=sub.sem)
            sems.append(sub.sem)
            last = i == self.mb - 1
            stream = main if last else self._side[i]
            if not last:
                self._a_done[i].record(main)
                stream.wait_event(self._a_done[i])
            with torch.cuda.stream(stream):
                sub._depth_ready, sub._pmap_ready = _events
                try:
                    sub._glue_mapper_window(rgb[sl], depth[sl], pose_delta[sl], local_map[sl], poses[sl], partial_map[sl],
                                            map_out=out_map[sl], fp_out=self.fp_out[sl])
                finally:
                    sub._depth_ready, sub._pmap_ready = None, None
                sub.pred.forward_device(partial_map[sl], apply_sigmoid=True, out=sub.pred_out)
                if not last:
                    self._rest_done[i].record(stream)
        for ev in self._rest_done:
            main.wait_event(ev)
        return sems, self.fp_out, out_map, poses, self.pred_out

    def step_host(self, rgb_h, depth_h, pose_delta_h, partial_map_h, local_map, poses):
        """End-to-end step with pinned HOST inputs / outputs, like PerceptionPipeline.step_host: only the frames gate
        Mask-RCNN, the other inputs are copied on a copy stream beside it."""
        d = self.device
        if self._dev_in is None:
            self._dev_in = (torch.empty_like(rgb_h, device=d), torch.empty_like(depth_h, device=d),
                            torch.empty_like(pose_delta_h, device=d), torch.empty_like(partial_map_h, device=d))
            self._host_out = (torch.empty(self.pred_out.shape, dtype=torch.float32).pin_memory(),
                              torch.empty((self.E, 3), dtype=torch.float32).pin_memory(),
                              torch.empty(self.fp_out.shape, dtype=torch.float32).pin_memory())
        main = torch.cuda.current_stream(d)
        self._dev_in[0].copy_(rgb_h, non_blocking=True)
        self._fence.record(main)
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(self._fence)
            self._dev_in[1].copy_(depth_h, non_blocking=True)
            self._dev_in[2].copy_(pose_delta_h, non_blocking=True)
            self._depth_ev.record(self._copy)
            self._dev_in[3].copy_(partial_map_h, non_blocking=True)
            self._pmap_ev.record(self._copy)
        rgb, depth, delta, pmap = self._dev_in
        _, fp, new_map, poses, pred = self.step_device(rgb, depth, delta, local_map, poses, pmap,
                                                       _events=(self._depth_ev, self._pmap_ev))
        self._host_out[1].copy_(poses, non_blocking=True)
        self._host_out[2].copy_(fp, non_blocking=True)
        self._host_out[0].copy_(pred, non_blocking=True)
        main.synchronize()
        return self._host_out[0], self._host_out[1], self._host_out[2], new_map

    def h2d_bytes(self, rgb_h, depth_h, pose_delta_h, partial_map_h):
        return sum(t.numel() * t.element_size() for t in (rgb_h, depth_h, pose_delta_h, partial_map_h))

    def d2h_bytes(self):
        return sum(t.numel() * t.element_size() for t in self._host_out) if self._host_out else 0
