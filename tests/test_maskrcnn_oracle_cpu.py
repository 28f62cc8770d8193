"""CPU: structural pins of the Mask-RCNN oracle (oracle/maskrcnn.py).  detectron2 is not installable here and the
reference holds no golden values for stage A (parity unpinned), so the restatement is checked
against the facts the config fixes: parameter count, stage shapes, resize rule, anchor census, and the
self-consistency of its discrete stages on a small frame."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import maskrcnn as O


def test_parameter_census_and_keys():
    w = O.synth_weights(0)
    n = sum(v.numel() for k, v in w.items() if not k.endswith(("running_mean", "running_var")))
    assert abs(n / 1e6 - 62.9) < 0.3, n  # R101-FPN Mask-RCNN with 9 classes (~63 M)
    assert w["backbone.bottom_up.stem.conv1.weight"].shape == (64, 3, 7, 7)
    assert w["backbone.bottom_up.res4.22.conv2.weight"].shape == (256, 256, 3, 3)
    assert "backbone.bottom_up.res4.1.shortcut.weight" not in w
    assert w["roi_heads.box_head.fc1.weight"].shape == (1024, 12544)
    assert w["roi_heads.box_predictor.bbox_pred.weight"].shape == (36, 1024)
    assert w["roi_heads.mask_head.deconv.weight"].shape == (256, 256, 2, 2)
    assert w["roi_heads.mask_head.predictor.weight"].shape == (9, 256, 1, 1)


def test_resize_rule_and_anchor_census():
    assert O.resized_shape(480, 640, O.Cfg()) == (800, 1067)          # yaml:28,30 on the 640 x 480 sensor
    assert O.resized_shape(480, 1920, O.Cfg()) == (333, 1333)         # max-size cap
    total = 0
    for l, s in enumerate(O.FPN_STRIDES):
        hh, ww = -(-800 // s), -(-1088 // s)
        a = O.grid_anchors(hh, ww, s, O.ANCHOR_SIZES[l])
        assert a.shape == (hh * ww * 3, 4)
        total += a.shape[0]
    assert total == 217413
    a = O.grid_anchors(2, 3, 4, 32)
    # order (y, x, anchor); first anchor of cell (0,1) is the ratio-0.5 one shifted by the stride
    assert torch.allclose(a[3], torch.tensor([4 - 22.627417, -11.313708, 4 + 22.627417, 11.313708]))
    assert torch.allclose(O.cell_anchors(32)[1], torch.tensor([-16., -16., 16., 16.]))


def test_small_frame_end_to_end_invariants():
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    w = O.synth_weights(0)
    img = O.synth_rgb(2, 120, 160)
    cfg = O.Cfg(min_size=192, max_size=333, score_thresh=0.3)
    taps = {}
    r = O.forward(img, w, cfg, taps=taps)
    assert taps["input"].shape == (1, 3, 192, 256) and taps["image_size"] == (192, 256)
    assert [tuple(taps["feats"][f"res{i}"].shape[1:]) for i in (2, 3, 4, 5)] == [(256, 48, 64), (512, 24, 32), (1024, 12, 16), (2048, 6, 8)]
    assert tuple(taps["pyr"]["p6"].shape[1:]) == (256, 3, 4)
    p, s = taps["proposals"], taps["proposal_logits"]
    assert 0 < p.shape[0] <= 1000 and (s[:-1] >= s[1:]).all()
    assert (p[:, 0] >= 0).all() and (p[:, 2] <= 256).all() and (p[:, 3] <= 192).all()
    assert ((p[:, 2] - p[:, 0]) > 0).all() and ((p[:, 3] - p[:, 1]) > 0).all()
    n = r["boxes"].shape[0]
    assert n <= 100 and r["masks"].shape == (n, 120, 160) and r["masks"].dtype == torch.bool
    assert (r["scores"] > 0.3).all() and (r["scores"][:-1] >= r["scores"][1:]).all()
    # per-class NMS really happened: no two kept detections of one class overlap by more than 0.5 (network coords)
    from torchvision.ops import box_iou
    b, c = taps["det_boxes"], taps["det_classes"]
    iou = box_iou(b, b) - torch.eye(len(b))
    assert not ((iou > 0.5) & (c[:, None] == c[None, :])).any()
    sem, bgr = O.get_prediction(img, w, cfg, sem_pred_prob_thr=0.3, goal_thr=0.3)
    assert sem.shape == (120, 160, 10) and (sem[:, :, 9] == 0).all() and np.array_equal(bgr, img[:, :, ::-1])
    assert np.array_equal(sem, O.accumulate(r["masks"], r["scores"], r["classes"], 9, 0.3, 0.3, None, 120, 160).numpy())


def test_bf16_emulation_is_close_to_fp32():
    w = O.synth_weights(0)
    img = O.synth_rgb(2, 120, 160)
    cfg = O.Cfg(min_size=192, max_size=333)
    with torch.no_grad():
        x, _ = O.preprocess(img, cfg)
        a = O.backbone(x, w)["res5"]
        b = O.backbone(x, w, emulate_bf16=True)["res5"]
    assert (a - b).abs().max() / a.abs().max() < 0.1


D2_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "maskrcnn_d2_*.npz")))


@pytest.mark.skipif(not D2_GOLDEN, reason="no detectron2 golden: tests/golden/make_maskrcnn_golden.py has to run where detectron2 0.6 "
                                          "imports (stage-A oracle stays parity-unpinned until then)")
@pytest.mark.parametrize("path", D2_GOLDEN or [None])
def test_oracle_matches_detectron2_golden(path):
    """The oracle against outputs of the UNMODIFIED reference wrapper on detectron2 (when the fixture exists)."""
    z = np.load(path)
    thr, goal_thr = float(z["thr"]), float(z["goal_thr"])
    goal_cat = None if int(z["goal_cat"]) < 0 else int(z["goal_cat"])
    frame = O.synth_rgb(int(z["seed"]))
    taps = {}
    ref = O.forward(frame, O.synth_weights(0), O.Cfg(score_thresh=thr), taps=taps)
    assert tuple(taps["input"].shape[2:]) == tuple(z["input_hw"]) and tuple(taps["image_size"]) == tuple(z["image_size"])
    for k in ("res2", "res3", "res4", "res5"):
        got = taps["feats"][k][0, :, ::8, ::8].numpy()
        assert np.abs(got - z[k].astype(np.float32)).max() <= 2e-3 * np.abs(z[k].astype(np.float32)).max() + 1e-3, k   # fp16 storage
        assert abs(float(taps["feats"][k].double().sum()) - float(z[k + "_sum"])) <= 1e-4 * abs(float(z[k + "_sum"])) + 1.0
    for k in ("p2", "p3", "p4", "p5", "p6"):
        got = taps["pyr"][k][0, :, ::8, ::8].numpy()
        assert np.abs(got - z[k].astype(np.float32)).max() <= 2e-3 * np.abs(z[k].astype(np.float32)).max() + 1e-3, k
    assert taps["proposals"].shape[0] == z["prop_boxes"].shape[0]
    assert np.abs(taps["proposals"].numpy() - z["prop_boxes"]).max() <= 2e-2          # pixels; same order
    assert taps["det_boxes"].shape[0] == z["det_boxes"].shape[0]
    assert np.array_equal(taps["det_classes"].numpy(), z["det_classes"])
    assert np.abs(taps["det_scores"].numpy() - z["det_scores"]).max() <= 1e-4
    sem = O.accumulate(ref["masks"], ref["scores"], ref["classes"], 9, thr, goal_thr, goal_cat, 480, 640).numpy()
    differing = int((sem != z["sem_u8"].astype(np.float32)).sum())
    assert differing <= 1e-5 * sem.size, f"{differing} category-stack cells differ from the reference"   # paste threshold at exactly 0.5
