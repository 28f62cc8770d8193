"""Stage C drop-in: the map-completion model behind the reference's own call surface.

Mirrors (same names, argument meaning, return types, error behaviour):
  * ``PEANUT_Prediction_Model(args)`` / ``.get_prediction(full_map)``   nav/agent/prediction.py:140-158
  * ``run_inference(model, full_map)``                                  nav/agent/prediction.py:112-137
    (itself a copy of ``mmseg.apis.inference_segmentor``, prediction/mmseg/apis/inference.py:70-99)
  * ``init_segmentor(config, checkpoint, device)``                      prediction/mmseg/apis/inference.py:12-40

The forward pass runs entirely in libpeanut_b200.so (tcgen05 convolutions + fused resize/sigmoid);
this module only parses the config, hands the checkpoint tensors to the library and moves buffers.
The reference's test pipeline (MapFromArray -> MultiScaleFlipAug(1.0, no flip) -> Resize -> ImageToTensor
-> Collect) is a numerical identity on the map and is therefore not re-enacted.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib


class Config(dict):
    """Minimal stand-in for ``mmcv.Config``: python-file configs, attribute access."""

    @staticmethod
    def fromfile(path):
        ns = {}
        with open(path, "r") as f:
            exec(compile(f.read(), path, "exec"), ns)  # same trust model as mmcv.Config.fromfile
        return Config({k: v for k, v in ns.items() if not k.startswith("_")})

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return Config(v) if isinstance(v, dict) and not isinstance(v, Config) else v


def _default_cfg(in_channels=14, num_classes=6):
    return Config(model=dict(type="EncoderDecoder",
                             backbone=dict(type="ResNetV1c", depth=50, in_channels=in_channels,
                                           strides=(1, 2, 1, 1), dilations=(1, 1, 2, 4), contract_dilation=True),
                             decode_head=dict(type="PSPHead", in_channels=2048, channels=512,
                                              pool_scales=(1, 2, 3, 6), num_classes=num_classes, align_corners=False),
                             test_cfg=dict(mode="whole")))


class Segmentor:
    """What ``init_segmentor`` returns: carries ``cfg`` / ``CLASSES`` like the mmseg model object."""

    def __init__(self, cfg, state_dict, device, precision="tf32", classes=None):
        m = cfg["model"]
        bb, head = m["backbone"], m["decode_head"]
        if bb.get("type", "ResNetV1c") != "ResNetV1c" or bb.get("depth", 50) != 50:
            raise NotImplementedError("peanut_b200 implements the reference's ResNetV1c-50 backbone only")
        if tuple(bb.get("strides", (1, 2, 1, 1))) != (1, 2, 1, 1) or tuple(bb.get("dilations", (1, 1, 2, 4))) != (1, 1, 2, 4):
            raise NotImplementedError("peanut_b200 implements strides (1,2,1,1) / dilations (1,1,2,4) only")
        if head.get("type", "PSPHead") != "PSPHead" or tuple(head.get("pool_scales", (1, 2, 3, 6))) != (1, 2, 3, 6):
            raise NotImplementedError("peanut_b200 implements PSPHead with pool scales (1,2,3,6) only")
        if m.get("test_cfg", {}).get("mode", "whole") != "whole":
            raise NotImplementedError("only whole-image inference is on the reference's path")
        self.cfg = cfg
        self.in_channels = int(bb.get("in_channels", 3))
        self.num_classes = int(head["num_classes"])
        self.CLASSES = classes
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("peanut_b200 has no CPU path: device must be a CUDA device")
        self.device = torch.device("cuda", dev.index if dev.index is not None else 0)
        self.precision = _lib.precision_code(precision)
        self.ctx = _lib.Context(self.device.index)
        self.ctx.set_weights({k: v for k, v in state_dict.items()
                              if not k.endswith("num_batches_tracked") and not k.startswith("auxiliary_head")})
        self._built = None
        self._pinned = {}

    # -- the reference calls .eval() and next(model.parameters()).device on the mmseg module
    def eval(self):
        return self

    def _ensure_built(self, B, C, H, W):
        key = (B, C, H, W)
        if self._built != key:
            if C != self.in_channels:
                raise RuntimeError(f"expected a {self.in_channels}-channel map, got {C}")
            _lib.check(self.ctx.lib.pn_prednet_build(self.ctx.handle, B, C, H, W, self.num_classes, self.precision))
            self._built = key

    def num_launches(self):
        return int(self.ctx.lib.pn_prednet_num_launches(self.ctx.handle))

    def forward_device(self, maps, apply_sigmoid=False, out=None):
        """maps: float32 CUDA tensor [B,C,H,W] -> float32 CUDA tensor [B,num_classes,H,W] (no host sync)."""
        if not (maps.is_cuda and maps.dtype == torch.float32 and maps.dim() == 4):
            raise TypeError("forward_device expects a float32 CUDA tensor [B,C,H,W]")
        maps = maps.contiguous()
        B, C, H, W = maps.shape
        self._ensure_built(B, C, H, W)
        if out is None:
            out = torch.empty((B, self.num_classes, H, W), dtype=torch.float32, device=maps.device)
        stream = torch.cuda.current_stream(maps.device).cuda_stream
        _lib.check(self.ctx.lib.pn_prednet_forward(self.ctx.handle, maps.data_ptr(), int(bool(apply_sigmoid)),
                                                   out.data_ptr(), ctypes.c_void_p(stream)))
        return out

    def forward_host(self, maps_host, apply_sigmoid=False, out_host=None):
        """maps_host: float32 host tensor [B,C,H,W] (pinned => async DMA) -> pinned float32 host tensor."""
        B, C, H, W = maps_host.shape
        self._ensure_built(B, C, H, W)
        if out_host is None:
            key = (B, self.num_classes, H, W)
            if key not in self._pinned:
                self._pinned[key] = torch.empty(key, dtype=torch.float32).pin_memory()
            out_host = self._pinned[key]
        _lib.check(self.ctx.lib.pn_prednet_forward_host(self.ctx.handle, maps_host.data_ptr(),
                                                        int(bool(apply_sigmoid)), out_host.data_ptr()))
        return out_host

    def profile(self, iters=5):
        """[(op name, milliseconds, FLOPs)] per recorded launch of the currently built network."""
        return _lib.net_profile(self.ctx, _lib.PN_NET_PREDNET, iters)

    def flops(self):
        """Algorithmic FLOPs of one forward of the currently built network (2*MAC over conv layers)."""
        out = ctypes.c_double()
        _lib.check(self.ctx.lib.pn_prednet_flops(self.ctx.handle, ctypes.byref(out)))
        return float(out.value)

    def read_tap(self, which):
        B, C, H, W = self._built
        d = lambda v: (v - 1) // 2 + 1  # one stride-2 stage (3x3 pad 1 / maxpool 3x3 pad 1)
        h, w = d(d(d(H))), d(d(d(W)))
        shape = (B, 2048, h, w) if which == 0 else (B, self.num_classes, h, w)
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        _lib.check(self.ctx.lib.pn_prednet_read_tap(self.ctx.handle, which, out.data_ptr(), None))
        torch.cuda.synchronize(self.device)
        return out


def init_segmentor(config, checkpoint=None, device="cuda:0", precision="tf32", state_dict=None):
    """prediction/mmseg/apis/inference.py:12-40.  ``checkpoint`` is an mmcv-format .pth
    ({'state_dict': ..., 'meta': {'CLASSES': ...}}); ``state_dict`` may be given directly instead."""
    if isinstance(config, str):
        config = Config.fromfile(config)
    elif not isinstance(config, dict):
        raise TypeError("config must be a filename or Config object, but got {}".format(type(config)))
    classes = None
    if checkpoint is not None:
        ckpt = torch.load(checkpoint, map_location="cpu", weights_only=False)
        state_dict = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
        classes = ckpt["meta"]["CLASSES"]  # KeyError if absent, as in the reference (inference.py:35)
    if state_dict is None:
        raise RuntimeError("init_segmentor: no checkpoint given (random initialisation is not supported)")
    return Segmentor(Config(config), state_dict, device, precision=precision, classes=classes)


def _as_host_batch(full_map):
    arr = np.asarray(full_map)
    if arr.ndim != 3:
        raise ValueError("full_map must be [C,H,W]")
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))[None]


def run_inference(model, full_map):
    """nav/agent/prediction.py:112-137: full_map ndarray [C,H,W] -> [ndarray [num_classes,H,W]] raw logits.
    Also accepts a CUDA tensor (then returns a list with one CUDA tensor, no host round trip)."""
    if torch.is_tensor(full_map) and full_map.is_cuda:
        return [model.forward_device(full_map[None].float(), apply_sigmoid=False)[0]]
    out = model.forward_host(_as_host_batch(full_map), apply_sigmoid=False)
    return [out[0].numpy().copy()]


class PEANUT_Prediction_Model():
    """nav/agent/prediction.py:140-158."""

    def __init__(self, args, state_dict=None, precision=None):
        self.args = args
        ckpt = getattr(args, "pred_model_wts", None)
        cfg_path = getattr(args, "pred_model_cfg", None)
        # the reference opens args.pred_model_cfg unconditionally (prediction.py:146) and raises if it is missing; only a
        # namespace WITHOUT the attribute (synthetic-weight callers) falls back to the reference's config values
        cfg = Config.fromfile(cfg_path) if cfg_path else _default_cfg()
        # fp32-parity path by default (tf32 operands, the tolerances of DESIGN section 3); "bf16" is the throughput opt-in
        precision = precision or getattr(args, "pn_precision", "tf32")
        device = ("cuda:" + str(args.sem_gpu_id)) if args else "cuda:0"
        self.model = init_segmentor(cfg, checkpoint=ckpt if state_dict is None else None, device=device,
                                    precision=precision, state_dict=state_dict)
        self.model.eval()
        self.model.cfg = cfg

    def get_prediction(self, full_map):
        """ndarray [C,H,W] -> ndarray [num_classes,H,W] probabilities (sigmoid fused on the device;
        the reference applies scipy.special.expit on the host, prediction.py:22-23,158)."""
        if torch.is_tensor(full_map) and full_map.is_cuda:
            return self.model.forward_device(full_map[None].float(), apply_sigmoid=True)[0]
        out = self.model.forward_host(_as_host_batch(full_map), apply_sigmoid=True)
        return out[0].numpy().copy()
