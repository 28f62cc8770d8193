"""Stage A drop-in: ``SemanticPredMaskRCNN`` behind the reference's own call surface.

Mirrors nav/agent/utils/segmentation.py:28-69: ``SemanticPredMaskRCNN(args)`` (needs ``args.sem_pred_prob_thr``,
``args.goal_thr``, ``args.seg_model_wts``, ``args.sem_gpu_id``), ``.n_cats``, and
``.get_prediction(img, depth=None, goal_cat=None) -> (float32 [H, W, n_cats + 1], img_bgr)`` with a host uint8 RGB
frame in and a fresh host array out (call site nav/agent/agent_helper.py:220-225).  The detectron2
``DefaultPredictor`` (resize, R101-FPN, RPN, ROI heads, mask paste) and the per-instance accumulation loop run
entirely in libpeanut_b200.so with no per-instance host synchronisation; this module only hands the checkpoint
tensors to the library and moves buffers.  ``compress_sem_map`` (segmentation.py:65-69) is host numpy in the
reference and stays host numpy here.
"""
import ctypes
import pickle

import numpy as np
import torch

from . import _lib

STAGES = ("preprocess", "backbone", "fpn", "rpn_head", "rpn_proposals", "box_head", "detections", "mask_head", "paste")


def load_detectron2_checkpoint(path):
    """detectron2 checkpoints: ``.pth`` = {'model': state_dict, ...}; model-zoo ``.pkl`` = {'model': {name: ndarray}}."""
    if str(path).endswith(".pkl"):
        with open(path, "rb") as f:
            ckpt = pickle.load(f, encoding="latin1")
    else:
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = ckpt["model"] if "model" in ckpt else ckpt
    return {k: (v if torch.is_tensor(v) else torch.from_numpy(np.asarray(v))) for k, v in sd.items()}


class MaskRCNN:
    """The batched device engine (B frames per call) the reference-facing class wraps."""

    def __init__(self, state_dict, device="cuda:0", precision="tf32", batch=1, height=480, width=640, cfg=None):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("peanut_b200 has no CPU path: device must be a CUDA device")
        self.device = torch.device("cuda", dev.index if dev.index is not None else 0)
        self.precision = _lib.precision_code(precision)
        self.cfg = cfg or default_cfg()
        self.n_cats = int(self.cfg.num_classes)
        self.B, self.H, self.W = int(batch), int(height), int(width)
        self.ctx = _lib.Context(self.device.index)
        self.ctx.set_weights({k: v for k, v in state_dict.items() if not k.startswith("pixel_")})
        _lib.check(self.ctx.lib.pn_maskrcnn_build(self.ctx.handle, self.B, self.H, self.W, self.precision,
                                                  ctypes.byref(self.cfg)))
        self.ctx.clear_weights()
        self._pinned = None

    def num_launches(self):
        return int(self.ctx.lib.pn_maskrcnn_num_launches(self.ctx.handle))

    def input_size(self):
        a, b = (ctypes.c_int * 2)(), (ctypes.c_int * 2)()
        _lib.check(self.ctx.lib.pn_maskrcnn_input_size(self.ctx.handle, a, b))
        return (a[0], a[1]), (b[0], b[1])

    def _check_rgb(self, rgb, cuda):
        if not (torch.is_tensor(rgb) and rgb.dtype == torch.uint8 and tuple(rgb.shape) == (self.B, self.H, self.W, 3)
                and rgb.is_cuda == cuda):
            raise TypeError(f"expected a uint8 {'CUDA' if cuda else 'host'} tensor [{self.B},{self.H},{self.W},3]")

    def forward_device(self, rgb, goal_cat=None, score_thresh=0.95, sem_pred_prob_thr=0.95, goal_thr=0.985, out=None):
        """rgb uint8 CUDA [B,H,W,3] (RGB) -> float32 CUDA [B,H,W,n_cats+1]; goal_cat: int32 CUDA [B] or None."""
        self._check_rgb(rgb, True)
        rgb = rgb.contiguous()
        if out is None:
            out = torch.empty((self.B, self.H, self.W, self.n_cats + 1), dtype=torch.float32, device=rgb.device)
        if goal_cat is not None and not (goal_cat.is_cuda and goal_cat.dtype == torch.int32 and goal_cat.numel() == self.B):
            raise TypeError("goal_cat must be an int32 CUDA tensor [B]")
        stream = torch.cuda.current_stream(rgb.device).cuda_stream
        _lib.check(self.ctx.lib.pn_maskrcnn_forward(self.ctx.handle, rgb.data_ptr(),
                                                    None if goal_cat is None else goal_cat.data_ptr(), score_thresh,
                                                    sem_pred_prob_thr, goal_thr, out.data_ptr(), ctypes.c_void_p(stream)))
        return out

    def forward_host(self, rgb_host, goal_cat=None, score_thresh=0.95, sem_pred_prob_thr=0.95, goal_thr=0.985, out_host=None):
        """rgb_host uint8 host tensor [B,H,W,3] (pinned => async DMA) -> pinned float32 host tensor."""
        self._check_rgb(rgb_host, False)
        if out_host is None:
            if self._pinned is None:
                self._pinned = torch.empty((self.B, self.H, self.W, self.n_cats + 1), dtype=torch.float32).pin_memory()
            out_host = self._pinned
        goal = None
        if goal_cat is not None:
            goal = np.ascontiguousarray(np.asarray(goal_cat, dtype=np.int32).reshape(self.B))
        _lib.check(self.ctx.lib.pn_maskrcnn_forward_host(self.ctx.handle, rgb_host.data_ptr(),
                                                         None if goal is None else goal.ctypes.data_as(ctypes.c_void_p),
                                                         score_thresh, sem_pred_prob_thr, goal_thr, out_host.data_ptr()))
        return out_host

    def profile(self, iters=3):
        return _lib.net_profile(self.ctx, _lib.PN_NET_MASKRCNN, iters)

    # ---- parity aids (tests only)
    def set_call(self, rgb, out, goal_cat=None, score_thresh=0.95, sem_pred_prob_thr=0.95, goal_thr=0.985):
        _lib.check(self.ctx.lib.pn_maskrcnn_set_call(self.ctx.handle, rgb.data_ptr(),
                                                     None if goal_cat is None else goal_cat.data_ptr(), score_thresh,
                                                     sem_pred_prob_thr, goal_thr, out.data_ptr(), None))

    def run_stages(self, first, end="end"):
        _lib.check(self.ctx.lib.pn_maskrcnn_run_stages(self.ctx.handle, first.encode(), end.encode(), None))

    def read_tap(self, name, shape, dtype=torch.float32):
        """Activations: shape = (B, C, H, W) -> NCHW float32.  Raw buffers: any shape / dtype, as stored."""
        out = torch.empty(shape, dtype=dtype, device=self.device)
        channels = shape[1] if (dtype == torch.float32 and len(shape) == 4) else 0
        _lib.check(self.ctx.lib.pn_maskrcnn_tap(self.ctx.handle, name.encode(), 0, channels, out.data_ptr(),
                                                out.numel() * out.element_size(), None))
        torch.cuda.synchronize(self.device)
        return out

    def write_tap(self, name, value, activation=False):
        value = value.to(self.device).contiguous()
        channels = value.shape[1] if activation else 0
        _lib.check(self.ctx.lib.pn_maskrcnn_tap(self.ctx.handle, name.encode(), 1, channels, value.data_ptr(),
                                                value.numel() * value.element_size(), None))
        torch.cuda.synchronize(self.device)


def default_cfg(**kw):
    """pn_maskrcnn_cfg with the values of nav/agent/utils/COCO-InstSeg/mask_rcnn_R_101_cat9.yaml."""
    c = _lib.MaskRcnnCfg(min_size_test=800, max_size_test=1333, rpn_pre_nms_topk=1000, rpn_post_nms_topk=1000,
                         rpn_nms_thresh=0.7, num_classes=9, box_nms_thresh=0.5, detections_per_image=100,
                         mask_threshold=0.5)
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


class SemanticPredMaskRCNN():
    """nav/agent/utils/segmentation.py:28-62."""

    def __init__(self, args, state_dict=None, precision=None):
        if state_dict is None:
            state_dict = load_detectron2_checkpoint(args.seg_model_wts)  # raises if missing, like DefaultPredictor
        # fp32-parity path by default (tf32 operands: the path the end-to-end detection tolerances are stated for);
        # args.pn_precision = "bf16" opts into the throughput path
        precision = precision or getattr(args, "pn_precision", "tf32")
        dev = args.sem_gpu_id if isinstance(args.sem_gpu_id, str) else "cuda:" + str(args.sem_gpu_id)
        self.engine = MaskRCNN(state_dict, device=dev, precision=precision, batch=1,
                               height=getattr(args, "env_frame_height", 480), width=getattr(args, "env_frame_width", 640))
        self.n_cats = self.engine.n_cats
        self.args = args
        self._rgb_pinned = torch.empty((1, self.engine.H, self.engine.W, 3), dtype=torch.uint8).pin_memory()

    def get_prediction(self, img, depth=None, goal_cat=None):
        args = self.args
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.shape != (self.engine.H, self.engine.W, 3):
            raise ValueError(f"expected a uint8 RGB frame of shape {(self.engine.H, self.engine.W, 3)}, got {img.dtype} {img.shape}")
        self._rgb_pinned[0].copy_(torch.from_numpy(np.ascontiguousarray(img)))
        out = self.engine.forward_host(self._rgb_pinned, None if goal_cat is None else [int(goal_cat)],
                                       score_thresh=args.sem_pred_prob_thr, sem_pred_prob_thr=args.sem_pred_prob_thr,
                                       goal_thr=args.goal_thr)
        return out[0].numpy().copy(), img[:, :, ::-1]


def compress_sem_map(sem_map):
    """nav/agent/utils/segmentation.py:65-69."""
    c_map = np.zeros((sem_map.shape[1], sem_map.shape[2]))
    for i in range(sem_map.shape[0]):
        c_map[sem_map[i] > 0.] = i + 1
    return c_map
