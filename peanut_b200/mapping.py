"""Stage B drop-in: ``Semantic_Mapping`` ("Sem_Map_Module") behind the reference's own call surface.

Mirrors nav/agent/mapping.py:10-179: an ``nn.Module`` constructed from the argparse namespace, on which the
caller invokes ``.to(device)`` / ``.eval()`` (nav/agent/agent_state.py:75-76) and
``forward(obs, pose_obs, maps_last, poses_last, agent_states)`` (agent_state.py:114-115, 273-274).
``poses_last`` is updated in place and also returned, ``map_pred`` is a fresh writable tensor - the
aliasing the reference's callers rely on.  All arithmetic runs in libpeanut_b200.so.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib


class Semantic_Mapping(nn.Module):

    def __init__(self, args, num_envs=1):
        super(Semantic_Mapping, self).__init__()
        self.args = args
        self.device = torch.device(args.device)
        if self.device.type != "cuda":
            raise RuntimeError("peanut_b200 has no CPU path: args.device must be a CUDA device")
        self.num_envs = int(num_envs)
        self.cfg = _lib.SemMapCfg(
            frame_height=args.frame_height, frame_width=args.frame_width, map_resolution=args.map_resolution,
            map_size_cm=args.map_size_cm, global_downscaling=args.global_downscaling, vision_range=args.vision_range,
            du_scale=args.du_scale, num_sem_categories=args.num_sem_categories, hfov=args.hfov,
            camera_height=args.camera_height, cat_pred_threshold=args.cat_pred_threshold,
            exp_pred_threshold=args.exp_pred_threshold, map_pred_threshold=args.map_pred_threshold)
        self.channels = 4 + args.num_sem_categories
        self.vr = args.vision_range
        self.cells = (args.map_size_cm // args.global_downscaling) // args.map_resolution
        self.h, self.w = args.frame_height, args.frame_width
        self.ctx = _lib.Context(self.device.index if self.device.index is not None else 0)
        _lib.check(self.ctx.lib.pn_semmap_build(self.ctx.handle, self.num_envs, ctypes.byref(self.cfg)))

    def _check(self, t, shape, name):
        if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == tuple(shape)):
            raise TypeError(f"{name}: expected float32 CUDA tensor of shape {tuple(shape)}, got "
                            f"{getattr(t, 'dtype', None)} {tuple(getattr(t, 'shape', ()))}")

    def forward_batch(self, obs, pose_obs, maps_last, poses_last, out=None, fp_out=None):
        """E environments at once: obs [E,C,h,w], pose_obs [E,3], maps_last [E,C,n,n] (any strides with unit
        x stride), poses_last [E,3] contiguous (updated in place).  Returns (fp_map_pred [E,vr,vr],
        map_pred [E,C,n,n], poses_last).  ``out`` / ``fp_out``: optional preallocated contiguous result tensors (no
        allocation on the calling stream then); ``out`` must not alias ``maps_last``."""
        E = self.num_envs
        self._check(obs, (E, self.channels, self.h, self.w), "obs")
        self._check(pose_obs, (E, 3), "pose_obs")
        self._check(maps_last, (E, self.channels, self.cells, self.cells), "maps_last")
        self._check(poses_last, (E, 3), "poses_last")
        if not poses_last.is_contiguous():
            raise ValueError("poses_last must be contiguous (it is updated in place)")
        obs = obs.contiguous()
        pose_obs = pose_obs.contiguous()
        if maps_last.stride(3) != 1:
            maps_last = maps_last.contiguous()
        strides = (ctypes.c_int64 * 3)(maps_last.stride(0), maps_last.stride(1), maps_last.stride(2))
        fp = fp_out if fp_out is not None else torch.empty((E, self.vr, self.vr), dtype=torch.float32, device=obs.device)
        if out is None:
            out = torch.empty((E, self.channels, self.cells, self.cells), dtype=torch.float32, device=obs.device)
        self._check(fp, (E, self.vr, self.vr), "fp_out")
        self._check(out, (E, self.channels, self.cells, self.cells), "out")
        if not (fp.is_contiguous() and out.is_contiguous()):
            raise ValueError("out / fp_out must be contiguous")
        stream = torch.cuda.current_stream(obs.device).cuda_stream
        _lib.check(self.ctx.lib.pn_semmap_forward(self.ctx.handle, obs.data_ptr(), pose_obs.data_ptr(),
                                                  maps_last.data_ptr(), strides, poses_last.data_ptr(), fp.data_ptr(),
                                                  out.data_ptr(), ctypes.c_void_p(stream)))
        return fp, out, poses_last

    def forward(self, obs, pose_obs, maps_last, poses_last, agent_states=None):
        """Reference signature (one environment): obs [1,C,h,w], pose_obs [3], maps_last [C,n,n], poses_last [3]
        -> (fp_map_pred [1,vr,vr], map_pred [C,n,n], pose_pred [3], current_poses [3])."""
        if self.num_envs != 1:
            raise RuntimeError("forward() is the single-environment reference signature; use forward_batch()")
        fp, out, _ = self.forward_batch(obs, pose_obs[None, :], maps_last[None, :], poses_last[None, :])
        return fp, out[0], poses_last, poses_last

    def read_ego(self):
        """Parity tap: (ego map [E, 2+S, vr, vr], stair-mask flags [E])."""
        ego = torch.empty((self.num_envs, self.channels - 2, self.vr, self.vr), dtype=torch.float32, device=self.device)
        flags = torch.empty((self.num_envs,), dtype=torch.int32, device=self.device)
        _lib.check(self.ctx.lib.pn_semmap_read_ego(self.ctx.handle, ego.data_ptr(), flags.data_ptr(), None))
        torch.cuda.synchronize(self.device)
        return ego, flags
