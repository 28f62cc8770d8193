"""Stage A parity on the GPU: CUDA Mask-RCNN (through the C-ABI and the reference-facing shim) against the CPU
oracle (oracle/maskrcnn.py) on the same seeded weights and frames, stage by stage.

The discrete stages (top-k, NMS, score filter, paste threshold) are checked EXACTLY by feeding them the oracle's
own inputs through the parity taps; the tensor-core stages are checked within stated floating-point tolerances:
  * tf32 path vs fp32 oracle: 5e-3 of the tensor's range per stage (10-bit operand mantissas, ~100 layers);
  * bf16 path vs the bf16-storage emulation of the oracle: 4e-2 of the range.
"""
import types

import numpy as np
import pytest
import torch

from oracle import maskrcnn as O
from peanut_b200 import segmentation as S

pytestmark = pytest.mark.gpu

H, W = 240, 320          # small camera frame for the stage tests (the e2e test uses the reference's 480 x 640)
MIN_SIZE, MAX_SIZE = 400, 667
THR = 0.3                # SCORE_THRESH_TEST for random weights (0.95 leaves too few detections to test NMS)


@pytest.fixture(scope="module")
def weights():
    return O.synth_weights(0)


@pytest.fixture(scope="module")
def frame():
    return O.synth_rgb(3, H, W)


@pytest.fixture(scope="module")
def ocfg():
    return O.Cfg(min_size=MIN_SIZE, max_size=MAX_SIZE, score_thresh=THR)


@pytest.fixture(scope="module")
def otaps(weights, frame, ocfg):
    taps = {}
    taps["result"] = O.forward(frame, weights, ocfg, taps=taps)
    return taps


def _engine(weights, precision, batch=1, h=H, w=W, min_size=MIN_SIZE, max_size=MAX_SIZE):
    return S.MaskRCNN(weights, precision=precision, batch=batch, height=h, width=w,
                      cfg=S.default_cfg(min_size_test=min_size, max_size_test=max_size))


@pytest.fixture(scope="module")
def eng_tf32(weights, frame):
    e = _engine(weights, "tf32")
    e._rgb = torch.from_numpy(frame)[None].cuda()
    e._out = torch.zeros((1, H, W, 10), device="cuda")
    e.set_call(e._rgb, e._out, score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR)
    return e


def _rel(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def test_resize_is_pil_exact_and_input_normalised(eng_tf32, frame, ocfg, otaps):
    e = eng_tf32
    e.run_stages("preprocess", "backbone")
    (hn, wn), (hp, wp) = e.input_size()
    assert (hn, wn) == O.resized_shape(H, W, ocfg) and (hp, wp) == tuple(otaps["input"].shape[2:])
    from PIL import Image
    ref = np.asarray(Image.fromarray(np.ascontiguousarray(frame[:, :, ::-1])).resize((wn, hn), Image.BILINEAR))
    got = e.read_tap("resized_u8", (1, hn, wn, 3), torch.uint8).cpu().numpy()[0]
    assert np.array_equal(got, ref), f"{(got != ref).sum()} resized pixels differ from PIL"
    # the stem input packs the horizontal taps of the stride-2 7x7 stem into channels, two output pixels (xo = 2j, 2j + 1) per
    # vector:  v[b, s*3 + c, y, j] = input[b, c, y, 4*j - 3 + s], s = 0..8  (0 outside), channels 27..31 zero - and two image
    # rows per stored pixel:  stored[b, h*32 + k, q, j] = v[b, k, 2q - 1 + h, j]  (row -1 and row hp are zeros)
    st = e.read_tap("stem_in", (1, 64, hp // 2 + 1, wp // 4)).cpu()
    assert (st[:, :32, 0] == 0).all() and (st[:, 32:, hp // 2] == 0).all()
    t = torch.zeros((1, 32, hp, wp // 4))
    t[:, :, 0::2] = st[:, 32:, :hp // 2]      # even rows y = 2q  (h = 1)
    t[:, :, 1::2] = st[:, :32, 1:]            # odd rows  y = 2q - 1  (h = 0), q = 1 .. hp / 2
    ref = torch.nn.functional.pad(otaps["input"], (3, 5))            # [1, 3, hp, wp + 8]
    for s in range(9):
        want = ref[:, :, :, s:s + wp:4]                               # x = 4*j - 3 + s  <=>  padded index 4*j + s
        # stored as tf32 (10-bit mantissa, round to nearest): |err| <= 2^-11 * 128
        assert (t[:, 3 * s:3 * s + 3] - want).abs().max().item() <= 0.0626, f"tap {s}"
    assert (t[:, 27:] == 0).all() and (t[:, :, hn:, :] == 0).all()


def test_backbone_fpn_rpn_head_tf32(eng_tf32, otaps):
    e = eng_tf32
    e.run_stages("preprocess", "rpn_proposals")
    for k, c in (("res2", 256), ("res3", 512), ("res4", 1024), ("res5", 2048)):
        ref = otaps["feats"][k]
        got = e.read_tap(k, tuple(ref.shape)).cpu()
        assert _rel(got, ref) <= 5e-3, k
    for k in ("p2", "p3", "p4", "p5", "p6"):
        ref = otaps["pyr"][k]
        got = e.read_tap(k, tuple(ref.shape)).cpu()
        assert _rel(got, ref) <= 5e-3, k
    for l, (obj, dlt) in enumerate(otaps["rpn"]):
        _, _, h, w = obj.shape
        got = e.read_tap(f"rpn_head.p{l + 2}", (1, h, w, 16)).cpu()
        ref = torch.cat((obj, dlt), 1).permute(0, 2, 3, 1)
        assert (got[..., :15] - ref).abs().max().item() <= 5e-3 * ref.abs().max().item(), f"rpn head p{l + 2}"


def _inject_rpn_heads(e, otaps):
    for l, (obj, dlt) in enumerate(otaps["rpn"]):
        buf = torch.zeros((1, obj.shape[2], obj.shape[3], 16))
        buf[..., :15] = torch.cat((obj, dlt), 1).permute(0, 2, 3, 1)
        e.write_tap(f"rpn_head.p{l + 2}", buf)


def test_rpn_proposals_exact_given_oracle_head(eng_tf32, otaps):
    """top-k per level, decode, clip, per-level NMS 0.7, merge to 1000: same survivors in the same order."""
    e = eng_tf32
    e.run_stages("preprocess", "rpn_proposals")
    _inject_rpn_heads(e, otaps)
    e.run_stages("rpn_proposals", "box_head")
    n = int(e.read_tap("prop_count", (1,), torch.int32).item())
    ref_b, ref_s = otaps["proposals"], otaps["proposal_logits"]
    assert n == ref_b.shape[0]
    got_s = e.read_tap("prop_scores", (1000,)).cpu()[:n]
    got_b = e.read_tap("prop_boxes", (1000, 4)).cpu()[:n]
    assert torch.equal(got_s, ref_s), "proposal scores (the selected anchors) differ"
    assert (got_b - ref_b).abs().max().item() <= 2e-3
    img = e.read_tap("prop_img", (1000,), torch.int32).cpu()
    assert (img[:n] == 0).all() and (img[n:] == -1).all()


def test_roi_align_and_box_head(eng_tf32, otaps, weights):
    e = eng_tf32
    e.run_stages("preprocess", "rpn_proposals")
    _inject_rpn_heads(e, otaps)
    e.run_stages("rpn_proposals", "detections")
    n = otaps["proposals"].shape[0]
    pyr = {k: e.read_tap(k, tuple(otaps["pyr"][k].shape)).cpu() for k in ("p2", "p3", "p4", "p5")}
    boxes = e.read_tap("prop_boxes", (1000, 4)).cpu()[:n]
    ref_pool = O.roi_pool(pyr, boxes, 7)
    got_pool = e.read_tap("box_pooled", (1000, 256, 7, 7)).cpu()[:n]
    # same inputs, same fp32 op order; the stored result is rounded to tf32 (2^-11 relative)
    assert (got_pool - ref_pool).abs().max().item() <= 6e-4 * ref_pool.abs().max().item()
    cls_ref, dlt_ref = O.box_head(got_pool, weights)
    out = e.read_tap("box_out", (1000, 64)).cpu()[:n]
    assert (out[:, :10] - cls_ref).abs().max().item() <= 5e-3 * cls_ref.abs().max().item()
    assert (out[:, 10:46] - dlt_ref).abs().max().item() <= 5e-3 * dlt_ref.abs().max().item()


def test_roi_align_extreme_boxes(eng_tf32, otaps):
    """ROIAlign on boxes the RPN never emits, to walk every path of k_roi_align: footprints 1..8 feature columns wide (the
    compile-time instances), wider than 8 (predicated batches), wider than 32 cells per bin (per-sample path), empty and
    out-of-image boxes.  Reference: torchvision's roi_align through the oracle's level assignment."""
    e = eng_tf32
    e.run_stages("preprocess", "rpn_proposals")
    _inject_rpn_heads(e, otaps)
    e.run_stages("rpn_proposals", "box_head")
    n = otaps["proposals"].shape[0]
    pyr = {k: e.read_tap(k, tuple(otaps["pyr"][k].shape)).cpu() for k in ("p2", "p3", "p4", "p5")}
    boxes = e.read_tap("prop_boxes", (1000, 4)).cpu()
    extreme = torch.tensor([
        [10.2, 10.7, 10.9, 11.1],            # a fraction of a cell
        [50.0, 50.0, 50.0, 50.0],            # empty: no samples, zeros
        [0.0, 0.0, 533.0, 400.0],            # the whole image
        [-50.0, -30.0, 600.0, 450.0],        # beyond every border
        [0.0, 100.0, 533.0, 108.0],          # 19 cells per bin wide on p2
        [100.0, 0.0, 108.0, 400.0],          # ... and tall
        [-2000.0, 100.0, 2500.0, 104.0],     # 80 cells per bin on p3: per-sample path, most samples outside the map
        [100.0, -1500.0, 104.0, 1500.0],
        [30.0, 40.0, 44.0, 230.0],           # one column wide, 27 rows per bin
        [5.0, 5.0, 228.0, 229.0], [5.0, 5.0, 120.0, 117.0], [300.0, 200.0, 532.9, 399.9],
    ])
    k = extreme.shape[0]
    assert n > k
    boxes[:k] = extreme
    e.write_tap("prop_boxes", boxes)
    e.run_stages("box_head", "detections")
    ref_pool = O.roi_pool(pyr, boxes[:n], 7)
    got_pool = e.read_tap("box_pooled", (1000, 256, 7, 7)).cpu()[:n]
    scale = ref_pool.abs().max().item()
    assert torch.isfinite(got_pool).all()
    assert (got_pool[:k] - ref_pool[:k]).abs().max().item() <= 6e-4 * scale
    assert (got_pool - ref_pool).abs().max().item() <= 6e-4 * scale
    assert got_pool[1].abs().max().item() == 0.0


def _inject_box_head(e, otaps):
    n = otaps["proposals"].shape[0]
    pb = torch.zeros((1000, 4))
    pb[:n] = otaps["proposals"]
    e.write_tap("prop_boxes", pb)
    e.write_tap("prop_count", torch.tensor([n], dtype=torch.int32))
    out = torch.zeros((1000, 64))
    out[:n, :10] = otaps["cls_logits"]
    out[:n, 10:46] = otaps["deltas"]
    e.write_tap("box_out", out)
    return n


def test_detections_exact_given_oracle_head(eng_tf32, otaps):
    """softmax, score filter, class-wise decode, per-class NMS 0.5, top 100: same detections in the same order."""
    e = eng_tf32
    e.run_stages("preprocess", "rpn_proposals")
    _inject_box_head(e, otaps)
    e.run_stages("detections", "mask_head")
    nd = int(e.read_tap("det_count", (1,), torch.int32).item())
    assert nd == otaps["det_boxes"].shape[0] and nd > 10
    cls = e.read_tap("det_classes", (100,), torch.int32).cpu()[:nd].long()
    sc = e.read_tap("det_scores", (100,)).cpu()[:nd]
    bx = e.read_tap("det_boxes", (100, 4)).cpu()[:nd]
    assert torch.equal(cls, otaps["det_classes"])
    assert (sc - otaps["det_scores"]).abs().max().item() <= 1e-6
    assert (bx - otaps["det_boxes"]).abs().max().item() <= 2e-3
    assert len(set(cls.tolist())) >= 2, "synthetic weights should exercise more than one class"


def test_mask_head_and_paste(eng_tf32, otaps, weights, ocfg):
    e = eng_tf32
    e.run_stages("preprocess", "rpn_proposals")
    _inject_box_head(e, otaps)
    e.run_stages("detections", "end")
    nd = otaps["det_boxes"].shape[0]
    pyr = {k: e.read_tap(k, tuple(otaps["pyr"][k].shape)).cpu() for k in ("p2", "p3", "p4", "p5")}
    det_boxes = e.read_tap("det_boxes", (100, 4)).cpu()[:nd]
    det_scores = e.read_tap("det_scores", (100,)).cpu()[:nd]
    classes = otaps["det_classes"]
    ref_pool = O.roi_pool(pyr, det_boxes, 14)
    got_pool = e.read_tap("mask_pooled", (100, 256, 14, 14)).cpu()[:nd]
    assert (got_pool - ref_pool).abs().max().item() <= 6e-4 * ref_pool.abs().max().item()
    # mask logits: [roi, py, px, dy, dx, 16] -> [roi, 9, 28, 28]
    raw = e.read_tap("mask_logits", (100, 14, 14, 2, 2, 16)).cpu()[:nd]
    got_logits = raw.permute(0, 5, 1, 3, 2, 4).reshape(nd, 16, 28, 28)[:, :9]
    ref_probs = O.mask_head(got_pool, classes, weights)
    got_probs = got_logits[torch.arange(nd), classes].sigmoid()
    assert (got_probs - ref_probs).abs().max().item() <= 5e-3
    # paste + accumulate: exact given the device's own mask probabilities and boxes
    b, s, c, masks = O.postprocess(det_boxes, det_scores, classes, got_probs, otaps["image_size"], H, W, ocfg)
    ref_sem = O.accumulate(masks, s, c, 9, THR, THR, None, H, W)
    got_sem = e._out.cpu()[0]
    diff = (got_sem != ref_sem)
    assert diff.float().mean().item() <= 1e-5, f"{int(diff.sum())} of {diff.numel()} cells differ"
    assert (got_sem[..., 9] == 0).all() and got_sem.sum() > 0


def test_bf16_features_vs_emulated_oracle(weights, frame, ocfg):
    e = _engine(weights, "bf16")
    rgb = torch.from_numpy(frame)[None].cuda()
    out = torch.zeros((1, H, W, 10), device="cuda")
    e.set_call(rgb, out, score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR)
    e.run_stages("preprocess", "rpn_proposals")
    with torch.no_grad():
        x, _ = O.preprocess(frame, ocfg)
        feats = O.backbone(x, weights, emulate_bf16=True)
        pyr = O.fpn(feats, weights, emulate_bf16=True)
    for k in ("res2", "res3", "res4", "res5"):
        assert _rel(e.read_tap(k, tuple(feats[k].shape)).cpu(), feats[k]) <= 4e-2, k
    for k in ("p2", "p3", "p4", "p5", "p6"):
        assert _rel(e.read_tap(k, tuple(pyr[k].shape)).cpu(), pyr[k]) <= 4e-2, k


def _match(boxes_a, cls_a, boxes_b, cls_b, iou_thr=0.9):
    from torchvision.ops import box_iou
    if boxes_a.numel() == 0 or boxes_b.numel() == 0:
        return 0
    iou = box_iou(boxes_a, boxes_b)
    iou[cls_a[:, None] != cls_b[None, :]] = 0
    return int((iou.max(1).values >= iou_thr).sum())


def test_end_to_end_reference_frame_tf32(weights):
    """480 x 640 frame through the whole pipeline at the reference's geometry (800 x 1067 -> 800 x 1088)."""
    frame = O.synth_rgb(11)
    cfg = O.Cfg(score_thresh=THR)
    taps = {}
    ref = O.forward(frame, weights, cfg, taps=taps)
    e = _engine(weights, "tf32", h=480, w=640, min_size=800, max_size=1333)
    assert e.input_size() == ((800, 1067), (800, 1088))
    sem = e.forward_device(torch.from_numpy(frame)[None].cuda(), score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR)
    torch.cuda.synchronize()
    nd = int(e.read_tap("det_count", (1,), torch.int32).item())
    bx = e.read_tap("det_boxes", (100, 4)).cpu()[:nd]
    cl = e.read_tap("det_classes", (100,), torch.int32).cpu()[:nd].long()
    n_ref = taps["det_boxes"].shape[0]
    assert n_ref > 0
    matched = _match(taps["det_boxes"], taps["det_classes"], bx, cl)
    # measured (tools/e2e_parity_probe.py, three frames): 90-94 of 100 detections matched, 99.64-99.73 % of the cells of the
    # category map agree.  The rest are instances whose score / NMS decision sits within the tf32 storage noise (2^-11 per
    # stored activation, amplified through 33 residual blocks to ~2e-3 of the feature range) of a threshold: random
    # weights and SCORE_THRESH 0.3 put ~100 candidates per frame there.
    assert matched >= 0.85 * n_ref, f"only {matched} of {n_ref} oracle detections found"
    ref_sem = O.accumulate(ref["masks"], ref["scores"], ref["classes"], 9, THR, THR, None, 480, 640)
    agree = ((sem.cpu()[0] > 0) == (ref_sem > 0)).float().mean().item()
    assert agree >= 0.99, f"category-mask agreement {agree}"


def test_end_to_end_reference_frame_fp32(weights):
    """Strict-parity mode end to end at the reference's geometry against the fp32 oracle: with fp32-level features the discrete
    decisions (top-k, NMS, score gate, paste threshold) fall the same way as the oracle's: every detection is found and the
    category stack is the oracle's up to the cells whose mask probability sits within the remaining noise of 0.5.  Measured on
    three frames (tools/fp32_parity_probe.py, profiles/r02_fp32_parity.txt): 100 of 100 detections matched at IoU >= 0.9 on
    every frame (tf32: 90-94), res5 features 4e-5 of their range (tf32: 2e-3), 99.944-99.987 % of the [480, 640, 10] stack
    bit-equal to the oracle's (tf32: 99.64-99.73 % agreeing as a binary mask)."""
    frame = O.synth_rgb(11)
    cfg = O.Cfg(score_thresh=THR)
    taps = {}
    ref = O.forward(frame, weights, cfg, taps=taps)
    e = _engine(weights, "fp32", h=480, w=640, min_size=800, max_size=1333)
    sem = e.forward_device(torch.from_numpy(frame)[None].cuda(), score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR)
    torch.cuda.synchronize()
    nd = int(e.read_tap("det_count", (1,), torch.int32).item())
    bx = e.read_tap("det_boxes", (100, 4)).cpu()[:nd]
    cl = e.read_tap("det_classes", (100,), torch.int32).cpu()[:nd].long()
    n_ref = taps["det_boxes"].shape[0]
    assert n_ref > 0
    matched = _match(taps["det_boxes"], taps["det_classes"], bx, cl)
    assert matched >= 0.98 * n_ref, f"only {matched} of {n_ref} oracle detections found"
    ref_sem = O.accumulate(ref["masks"], ref["scores"], ref["classes"], 9, THR, THR, None, 480, 640)
    got = sem.cpu()[0]
    equal = (got == ref_sem).float().mean().item()
    assert equal >= 0.9995, f"category stack equal on {equal} of the cells"


def test_end_to_end_reference_frame_bf16(weights):
    """The throughput path end to end at the reference's geometry, against the oracle with the same bf16 storage rounding
    emulated (oracle/maskrcnn.py `_r`): detections and the category stack.  bf16 storage noise (2^-8 per stored activation,
    ~1.5e-2 of the feature range at res5) flips far more near-threshold decisions than tf32 does - measured on three frames:
    54-63 of 100 detections matched at IoU >= 0.9, 98.3-98.5 % of the category-map cells agree (97.3-97.4 % bit-equal incl.
    overlap counts) - which is why the drop-in shims default to tf32 and bf16 is an explicit throughput opt-in."""
    frame = O.synth_rgb(11)
    cfg = O.Cfg(score_thresh=THR)
    taps = {}
    ref = O.forward(frame, weights, cfg, emulate_bf16=True, taps=taps)
    e = _engine(weights, "bf16", h=480, w=640, min_size=800, max_size=1333)
    sem = e.forward_device(torch.from_numpy(frame)[None].cuda(), score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR)
    torch.cuda.synchronize()
    nd = int(e.read_tap("det_count", (1,), torch.int32).item())
    bx = e.read_tap("det_boxes", (100, 4)).cpu()[:nd]
    cl = e.read_tap("det_classes", (100,), torch.int32).cpu()[:nd].long()
    n_ref = taps["det_boxes"].shape[0]
    assert n_ref > 0 and nd > 0
    matched = _match(taps["det_boxes"], taps["det_classes"], bx, cl)
    assert matched >= 0.40 * n_ref, f"only {matched} of {n_ref} oracle detections found"
    loose = _match(taps["det_boxes"], taps["det_classes"], bx, cl, iou_thr=0.5)
    assert loose >= 0.60 * n_ref, f"only {loose} of {n_ref} oracle detections found at IoU 0.5"
    ref_sem = O.accumulate(ref["masks"], ref["scores"], ref["classes"], 9, THR, THR, None, 480, 640)
    got = sem.cpu()[0]
    agree = ((got > 0) == (ref_sem > 0)).float().mean().item()
    assert agree >= 0.975, f"category-mask agreement {agree}"
    assert (got == got.round()).all() and float(got[..., 9].abs().sum()) == 0


def test_reference_api_and_goal_gate(weights):
    frame = O.synth_rgb(11)
    args = types.SimpleNamespace(sem_pred_prob_thr=THR, goal_thr=0.9999999, seg_model_wts=None, sem_gpu_id=0)
    model = S.SemanticPredMaskRCNN(args, state_dict=weights, precision="bf16")
    assert model.n_cats == 9
    sem, bgr = model.get_prediction(frame)
    assert sem.shape == (480, 640, 10) and sem.dtype == np.float32 and sem.flags.writeable
    assert np.array_equal(bgr, frame[:, :, ::-1])
    assert (sem == np.round(sem)).all() and sem.min() >= 0 and sem.sum() > 0
    present = [c for c in range(9) if sem[:, :, c].sum() > 0]
    goal = present[0]
    sem_g, _ = model.get_prediction(frame, goal_cat=goal)   # goal_thr ~ 1 removes (nearly) every goal-class instance
    assert sem_g[:, :, goal].sum() < sem[:, :, goal].sum()
    for c in present[1:]:
        assert np.array_equal(sem_g[:, :, c], sem[:, :, c])
    again, _ = model.get_prediction(frame)
    assert np.array_equal(again, sem), "CUDA-graph replay must be bit-stable"
    with pytest.raises(ValueError):
        model.get_prediction(frame[:100])


def test_batch_equals_single(weights):
    """Frames are independent: a frame's result does not depend on its slot in the batch (bit-exact, same engine), and a
    batch-1 engine agrees with a batch-2 engine up to the split-K choice (the K partition depends on the tile count, so
    fp32 sums are associated differently and a few threshold decisions may flip)."""
    frames = np.stack([O.synth_rgb(s, H, W) for s in (1, 2)])
    e2 = _engine(weights, "bf16", batch=2)
    both = e2.forward_device(torch.from_numpy(frames).cuda(), score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR).cpu()
    swapped = e2.forward_device(torch.from_numpy(frames[::-1].copy()).cuda(), score_thresh=THR, sem_pred_prob_thr=THR,
                                goal_thr=THR).cpu()
    assert torch.equal(swapped[1], both[0]) and torch.equal(swapped[0], both[1])
    e1 = _engine(weights, "bf16", batch=1)
    for i in range(2):
        one = e1.forward_device(torch.from_numpy(frames[i:i + 1]).cuda(), score_thresh=THR, sem_pred_prob_thr=THR,
                                goal_thr=THR).cpu()
        # with random weights and a low score threshold whole instances flip on 1-ulp feature differences
        agree = float((one[0] == both[i]).float().mean())
        assert agree >= 0.85, f"frame {i}: batch-1 and batch-2 engines agree on {agree:.4f} of the category-mask cells"
    assert both.sum() > 0


def test_batch_equals_single_fp32(weights):
    """The same comparison in the strict-parity mode, where the association order of the fp32 sums moves results by ~1e-6
    instead of a bf16 ulp: a batch-2 engine and a batch-1 engine (different tile counts, hence possibly different split-K
    partitions) then give the same category stack except for the odd cell on the paste threshold."""
    frames = np.stack([O.synth_rgb(s, H, W) for s in (1, 2)])
    e2 = _engine(weights, "fp32", batch=2)
    both = e2.forward_device(torch.from_numpy(frames).cuda(), score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR).cpu()
    n2 = e2.read_tap("det_count", (2,), torch.int32).cpu().tolist()
    e1 = _engine(weights, "fp32", batch=1)
    for i in range(2):
        one = e1.forward_device(torch.from_numpy(frames[i:i + 1]).cuda(), score_thresh=THR, sem_pred_prob_thr=THR,
                                goal_thr=THR).cpu()
        n1 = int(e1.read_tap("det_count", (1,), torch.int32).item())
        assert abs(n1 - n2[i]) <= 1
        agree = float((one[0] == both[i]).float().mean())
        assert agree >= 0.995, f"frame {i}: batch-1 and batch-2 engines agree on {agree:.5f} of the category-mask cells"
    assert both.sum() > 0


def test_no_detections_and_blank_frame(weights):
    """Edge cases of the device-side control flow: an unreachable score threshold leaves zero detections (the mask head's
    live-row count is 0, every tile is skipped) and the category stack is all zeros; a blank frame runs through the same
    launch list without NaNs; the engine then still produces the normal result for a normal frame (no stale state)."""
    e = _engine(weights, "bf16", batch=2)
    frames = np.stack([O.synth_rgb(s, H, W) for s in (5, 6)])
    normal = e.forward_device(torch.from_numpy(frames).cuda(), score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR).clone()
    assert normal.sum() > 0
    none = e.forward_device(torch.from_numpy(frames).cuda(), score_thresh=2.0, sem_pred_prob_thr=2.0, goal_thr=2.0).clone()
    torch.cuda.synchronize()
    assert float(none.abs().sum()) == 0.0
    assert e.read_tap("det_count", (2,), torch.int32).cpu().tolist() == [0, 0]
    blank = np.zeros_like(frames)
    blank[1] = 255
    out = e.forward_device(torch.from_numpy(blank).cuda(), score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR).clone()
    assert torch.isfinite(out).all() and float(out.min()) >= 0.0
    again = e.forward_device(torch.from_numpy(frames).cuda(), score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR)
    assert torch.equal(again, normal)
