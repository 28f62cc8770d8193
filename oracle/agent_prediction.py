"""Oracle for §8(f) N1a: ``Agent_State.update_prediction`` (nav/agent/agent_state.py:345-374) - the consumer of stage C.

TEST INFRASTRUCTURE ONLY.  Pinned: tests/golden/make_update_prediction_golden.py extracts the UNMODIFIED method source from
the reference file with ``ast`` (the module itself imports skimage / skfmm / habitat, absent here), executes it on a stub
state object and requires bit-equality with this restatement before it writes tests/golden/update_prediction.npz.

Restated semantics, line by line:
  :350-351  full_map[:, lmb0:lmb1, lmb2:lmb3] = local_map                   (in place)
  :354-355  window == full size: predict on the whole full map
  :357-364  else: centred crop [x1:x2, y1:y2] -> predict -> embed into a float64 zero canvas of the full size
  :369-371  target_pred = object_preds[goal_cat, lmb0:lmb1, lmb2:lmb3]
  :372      target_pred *= (local_map[1] < 0.5)                               (unexplored cells only)
The result is float64 when the crop branch ran (np.zeros canvas) and float32 (a view of the model output, multiplied in
place) otherwise.
"""
import numpy as np


def update_prediction(full_map, local_map, lmb, goal_cat, get_prediction, prediction_window):
    """full_map [C, W, H] and local_map [C, w, h] float32 arrays (full_map is updated in place); lmb = (r0, r1, c0, c1);
    get_prediction: callable([C, win, win] float32) -> [K, win, win] float32.  Returns target_pred [w, h]."""
    full_map[:, lmb[0]:lmb[1], lmb[2]:lmb[3]] = local_map
    full_w, full_h = full_map.shape[1], full_map.shape[2]
    if full_w == prediction_window and full_h == prediction_window:
        object_preds = get_prediction(full_map)
    else:
        x1 = full_w // 2 - prediction_window // 2
        x2 = x1 + prediction_window
        y1 = full_h // 2 - prediction_window // 2
        y2 = y1 + prediction_window
        object_preds = get_prediction(full_map[:, x1:x2, y1:y2])
        temp = np.zeros((object_preds.shape[0], full_w, full_h))
        temp[:, x1:x2, y1:y2] = object_preds
        object_preds = temp
    target_pred = object_preds[goal_cat, lmb[0]:lmb[1], lmb[2]:lmb[3]]
    target_pred *= local_map[1] < 0.5
    return target_pred


def synth_state(seed, full=96, local=48, channels=14):
    """Small synthetic agent state: (full_map, local_map, lmb) with the local window somewhere inside the full map."""
    rng = np.random.default_rng(seed)
    full_map = (rng.random((channels, full, full)) < 0.15).astype(np.float32) * rng.random((channels, full, full)).astype(np.float32)
    local_map = (rng.random((channels, local, local)) < 0.3).astype(np.float32) * rng.random((channels, local, local)).astype(np.float32)
    local_map[1] = (rng.random((local, local)) < 0.5).astype(np.float32) * rng.random((local, local)).astype(np.float32)
    r0 = int(rng.integers(0, full - local + 1))
    c0 = int(rng.integers(0, full - local + 1))
    return full_map, local_map, np.array([r0, r0 + local, c0, c0 + local])


def fake_prediction(x, num_classes=6):
    """Deterministic stand-in for the prediction model (fp32, values in (0, 1)): depends on every input channel."""
    x = np.asarray(x, np.float32)
    w = np.linspace(0.3, 1.7, x.shape[0], dtype=np.float32)[:, None, None]
    base = (x * w).sum(0, dtype=np.float32)
    out = np.stack([1.0 / (1.0 + np.exp(-(base * np.float32(0.5 + 0.25 * k) - np.float32(0.2 * k)))) for k in range(num_classes)])
    return out.astype(np.float32)
