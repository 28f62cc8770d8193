"""Device-side ``Agent_State.update_prediction`` (nav/agent/agent_state.py:345-374) - SURVEY.md section 8(f), N1a.

The reference stamps the local map into the full map, copies the prediction window of the full map to the HOST, runs the
map-completion net through ``get_prediction`` (host -> device -> host, all six class planes), embeds the result into a
full-size float64 canvas on the host and finally keeps one class plane inside the local-map bounds, masked to unexplored
cells.  Here the window never leaves the device: stamp and crop are device copies, the net runs on the cropped view and
one small kernel (``pn_target_pred``) produces the masked goal-category plane, so a single [local_w, local_h] plane
(0.9 MB at the reference geometry instead of 12.4 MB) is all that may travel to the host-side planner.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def update_prediction(full_map, local_map, lmb, goal_cat, model, prediction_window, as_numpy=True, object_preds=None):
    """full_map [C, W, H] and local_map [C, w, h]: float32 CUDA tensors (full_map is updated in place, like the reference's
    ``self.full_map``); lmb = (r0, r1, c0, c1) local-map bounds inside the full map; model: PEANUT_Prediction_Model (or its
    Segmentor); object_preds: optional precomputed [K, window, window] CUDA tensor of class probabilities (skips the net).

    Returns ``target_pred`` [w, h]: a host numpy array with the reference's dtype (float64 when the window is a crop of the
    full map, float32 when it is the whole map) or, with as_numpy=False, the float32 CUDA tensor."""
    if not (full_map.is_cuda and local_map.is_cuda and full_map.dtype == torch.float32 and local_map.dtype == torch.float32):
        raise TypeError("update_prediction expects float32 CUDA tensors")
    r0, r1, c0, c1 = (int(v) for v in lmb)
    C, full_w, full_h = full_map.shape
    lw, lh = local_map.shape[1], local_map.shape[2]
    if (r1 - r0, c1 - c0) != (lw, lh) or local_map.shape[0] != C:
        raise ValueError("local-map bounds do not match the local map")
    full_map[:, r0:r1, c0:c1] = local_map                                     # agent_state.py:350-351
    win = int(prediction_window)
    whole = full_w == win and full_h == win
    x1 = 0 if whole else full_w // 2 - win // 2                               # agent_state.py:357-360
    y1 = 0 if whole else full_h // 2 - win // 2
    seg = getattr(model, "model", model)
    if object_preds is None:
        window = full_map if whole else full_map[:, x1:x1 + win, y1:y1 + win]
        object_preds = seg.forward_device(window[None], apply_sigmoid=True)[0]   # get_prediction: expit(logits)
    if not (object_preds.is_cuda and object_preds.dtype == torch.float32 and tuple(object_preds.shape[1:]) == (win, win)):
        raise TypeError("object_preds must be a float32 CUDA tensor [K, window, window]")
    object_preds = object_preds.contiguous()
    explored = local_map[1]
    if explored.stride(1) != 1:
        explored = explored.contiguous()
    out = torch.empty((lw, lh), dtype=torch.float32, device=full_map.device)
    ctx = seg.ctx
    stream = torch.cuda.current_stream(full_map.device).cuda_stream
    _lib.check(ctx.lib.pn_target_pred(ctx.handle, object_preds.data_ptr(), int(object_preds.shape[0]), win, x1, y1, int(goal_cat),
                                      r0, c0, lw, lh, explored.data_ptr(), int(explored.stride(0)), out.data_ptr(),
                                      ctypes.c_void_p(stream)))
    if not as_numpy:
        return out
    host = out.cpu().numpy()
    return host if whole else host.astype(np.float64)
