"""ctypes binding of libpeanut_b200.so (the C-ABI declared in include/peanut_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
torch is used only for device memory and streams; no torch type crosses the ABI (raw pointers do).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpeanut_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

PN_BF16 = 0
PN_TF32 = 1
PN_FP32 = 2



def precision_code(precision):
    """"tf32" (fp32 storage, tf32 tensor-core operands, fp32 accumulate: the default of the drop-in shims), "bf16" (throughput
    path) or "fp32" (strict parity: every convolution as three tf32 tensor-core products of hi / lo operand halves, no rounding
    of stored activations - fp32-level accuracy at about a third of the tf32 path's speed)."""
    try:
        return {"bf16": PN_BF16, "tf32": PN_TF32, "fp32": PN_FP32}[precision]
    except KeyError:
        raise ValueError(f"precision must be 'tf32', 'bf16' or 'fp32', got {precision!r}") from None


_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_i64_p = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); must list every symbol include/peanut_b200.h declares.
PROTOTYPES = {
    "pn_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "pn_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pn_last_error": (ctypes.c_char_p, []),
    "pn_abi_version": (ctypes.c_int, []),
    "pn_set_weight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, _c_i64_p]),
    "pn_clear_weights": (ctypes.c_int, [ctypes.c_void_p]),
    "pn_prednet_build": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int] * 6),
    "pn_prednet_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pn_prednet_forward_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "pn_prednet_num_launches": (ctypes.c_int, [ctypes.c_void_p]),
    "pn_prednet_read_tap": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "pn_prednet_flops": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pn_net_num_ops": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "pn_net_profile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_int]),
    "pn_maskrcnn_build": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "pn_maskrcnn_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float,
                                           ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "pn_maskrcnn_forward_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float,
                                                ctypes.c_float, ctypes.c_void_p]),
    "pn_maskrcnn_num_launches": (ctypes.c_int, [ctypes.c_void_p]),
    "pn_maskrcnn_input_size": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pn_maskrcnn_set_call": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float,
                                            ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "pn_maskrcnn_run_stages": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p]),
    "pn_maskrcnn_tap": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_void_p]),
    "pn_pil_bilinear_coeffs": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pn_make_obs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6 +
                    [ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "pn_target_pred": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 9 +
                       [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "pn_map_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "pn_map_stamp_initial": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "pn_map_update_local": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "pn_map_update_full": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "pn_map_stamp_local": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] * 6 + [ctypes.c_void_p]),
    "pn_map_crop_window": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 9 + [ctypes.c_void_p, ctypes.c_int,
                                                                                                  ctypes.c_void_p]),
    "pn_map_quantize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p]),
    "pn_map_sample": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p] * 4),
    "pn_global_goal": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "pn_goal_map": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 3 + [ctypes.c_int] +
                    [ctypes.c_void_p] * 3),
    "pn_semmap_build": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "pn_semmap_forward": (ctypes.c_int, [ctypes.c_void_p] * 9),
    "pn_semmap_read_ego": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "pn_semmap_num_launches": (ctypes.c_int, [ctypes.c_void_p]),
    "pn_gather_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                        ctypes.POINTER(ctypes.c_void_p)]),
    "pn_gather_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pn_gather_connect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pn_gather_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pn_gather_result": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64)]),
    "pn_gather_status": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]),
    "pn_gather_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pn_conv_tuning_import": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_void_p]),
    "pn_conv_tuning_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "pn_conv_tuning_clear": (ctypes.c_int, []),
    "pn_conv_bench": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int] * 14 + [ctypes.c_void_p, ctypes.c_void_p]),
    "pn_conv2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 4 +
                  [ctypes.c_void_p] * 4 + [ctypes.c_int] * 8 + [ctypes.c_void_p]),
}



class SemMapCfg(ctypes.Structure):
    """struct pn_semmap_cfg"""
    _fields_ = [("frame_height", ctypes.c_int), ("frame_width", ctypes.c_int), ("map_resolution", ctypes.c_int),
                ("map_size_cm", ctypes.c_int), ("global_downscaling", ctypes.c_int), ("vision_range", ctypes.c_int),
                ("du_scale", ctypes.c_int), ("num_sem_categories", ctypes.c_int), ("hfov", ctypes.c_double),
                ("camera_height", ctypes.c_double), ("cat_pred_threshold", ctypes.c_float),
                ("exp_pred_threshold", ctypes.c_float), ("map_pred_threshold", ctypes.c_float)]


class MapCfg(ctypes.Structure):
    """struct pn_map_cfg"""
    _fields_ = [("num_channels", ctypes.c_int), ("full_w", ctypes.c_int), ("full_h", ctypes.c_int), ("local_w", ctypes.c_int),
                ("local_h", ctypes.c_int), ("map_resolution", ctypes.c_int), ("map_size_cm", ctypes.c_int),
                ("global_downscaling", ctypes.c_int), ("grid_resolution", ctypes.c_int), ("col_rad", ctypes.c_int),
                ("goal_reached_dist", ctypes.c_float), ("f64_cells", ctypes.c_int)]


class MapArrays(ctypes.Structure):
    """struct pn_map_arrays (device pointers)"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("full_map", "local_map", "full_pose", "local_pose", "origins", "lmb",
                                               "planner_pose_inputs", "loc", "dist_to_goal", "global_goal")]


class GoalCfg(ctypes.Structure):
    """struct pn_goal_cfg"""
    _fields_ = [("num_channels", ctypes.c_int), ("full_w", ctypes.c_int), ("full_h", ctypes.c_int), ("local_w", ctypes.c_int),
                ("local_h", ctypes.c_int), ("col_rad", ctypes.c_int), ("map_resolution", ctypes.c_int),
                ("dist_weight_temperature", ctypes.c_double)]


class GoalArrays(ctypes.Structure):
    """struct pn_goal_arrays (device pointers)"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("full_map", "collision_map", "visited_vis", "lmb", "loc", "target_pred", "dd", "dd_wt",
                                               "dd_wt_valid", "value", "global_goal", "goal_kind", "last_global_goal", "last_kind")]


class MaskRcnnCfg(ctypes.Structure):
    """struct pn_maskrcnn_cfg"""
    _fields_ = [("min_size_test", ctypes.c_int), ("max_size_test", ctypes.c_int), ("rpn_pre_nms_topk", ctypes.c_int),
                ("rpn_post_nms_topk", ctypes.c_int), ("rpn_nms_thresh", ctypes.c_float), ("num_classes", ctypes.c_int),
                ("box_nms_thresh", ctypes.c_float), ("detections_per_image", ctypes.c_int), ("mask_threshold", ctypes.c_float)]


PN_NET_PREDNET = 0
PN_NET_MASKRCNN = 1


def net_profile(ctx, which, iters=5):
    """[(op name, milliseconds, algorithmic FLOPs)] per recorded launch of a built network."""
    lib = ctx.lib
    n = int(lib.pn_net_num_ops(ctx.handle, which))
    if n < 0:
        raise RuntimeError("peanut_b200: network not built")
    ms = (ctypes.c_float * n)()
    fl = (ctypes.c_double * n)()
    names = ctypes.create_string_buffer(96 * n)
    check(lib.pn_net_profile(ctx.handle, which, iters, ms, fl, n, names, len(names)))
    return list(zip(names.value.decode().strip().split("\n"), [float(v) for v in ms], [float(v) for v in fl]))


_lib = None


def build(verbose=False):
    """Compile the CUDA sources in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libpeanut_b200.so failed")
    return LIB_PATH


def load():
    """dlopen the library and bind every exported symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(peanut_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    _load_default_tuning(lib)
    return lib


TUNING_DIR = os.path.join(_HERE, "tuning")


def _load_default_tuning(lib):
    """Import the committed launch-configuration tables (peanut_b200/tuning/*.txt) and $PN_CONV_TUNING_FILE, so that every
    process - in particular every rank of one job - builds identical networks without timing anything for the shapes the
    tables cover.  PN_CONV_TUNING_FILE=none skips the committed tables."""
    override = os.environ.get("PN_CONV_TUNING_FILE", "")
    paths = []
    if override != "none" and os.path.isdir(TUNING_DIR):
        paths += sorted(os.path.join(TUNING_DIR, f) for f in os.listdir(TUNING_DIR) if f.endswith(".txt"))
    if override and override != "none":
        if not os.path.exists(override):
            raise RuntimeError(f"PN_CONV_TUNING_FILE={override} does not exist")
        paths.append(override)
    for p in paths:
        with open(p, "rb") as f:
            tuning_import(f.read(), lib)


def tuning_import(text, lib=None):
    """text: bytes / str in the format of tuning_export(); returns the number of entries read."""
    lib = lib or load()
    if isinstance(text, str):
        text = text.encode()
    n = ctypes.c_int(0)
    st = lib.pn_conv_tuning_import(text, ctypes.byref(n))
    if st != 0:
        raise RuntimeError("peanut_b200: " + lib.pn_last_error().decode("utf-8", "replace"))
    return n.value


def tuning_export():
    """The process-wide launch-configuration table as text (one "key bn splits pair opt" line per conv shape)."""
    lib = load()
    need = ctypes.c_int64(0)
    check(lib.pn_conv_tuning_export(None, 0, ctypes.byref(need)))
    buf = ctypes.create_string_buffer(need.value + 16)
    check(lib.pn_conv_tuning_export(buf, len(buf), None))
    return buf.value.decode()


def check(status):
    if status != 0:
        raise RuntimeError("peanut_b200: " + load().pn_last_error().decode("utf-8", "replace"))


class Context:
    """One pn_ctx: a device, a weight store and the built stage networks."""

    def __init__(self, device=0):
        self.lib = load()
        h = ctypes.c_void_p()
        check(self.lib.pn_create(int(device), ctypes.byref(h)))
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pn_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, state_dict):
        """state_dict: name -> array-like (torch tensor or numpy), passed to the library as fp32."""
        for name, value in state_dict.items():
            if hasattr(value, "detach"):
                value = value.detach().cpu().numpy()
            arr = np.ascontiguousarray(value, dtype=np.float32)
            shape = (ctypes.c_int64 * max(arr.ndim, 1))(*arr.shape)
            check(self.lib.pn_set_weight(self.handle, name.encode(), arr.ctypes.data_as(ctypes.c_void_p), arr.ndim, shape))

    def clear_weights(self):
        check(self.lib.pn_clear_weights(self.handle))
