"""Independent cross-check of the stage-A oracle (oracle/maskrcnn.py).

detectron2 0.6 - the library the reference calls (nav/agent/utils/segmentation.py:30-45) - cannot be installed here, so the
oracle is a restatement that no reference-held vector pins (tests/golden/make_maskrcnn_golden.py is the script that writes
such vectors where detectron2 imports).  Until then every sub-function that has a counterpart in torchvision's detection
package - an independent implementation of the same published algorithms (Faster/Mask R-CNN, FPN) by different authors -
is checked against it on random inputs, with the KNOWN differences between the two libraries written down and neutralised:

  oracle function        torchvision counterpart                     known difference (how the test neutralises it)
  apply_deltas           _utils.BoxCoder.decode_single               none (same clamp log(1000/16))
  assign_levels          ops.poolers.LevelMapper                     eps: tv floor(k0 + log2(s/s0) + 1e-6), d2 floor(k0 + log2(s/s0 + 1e-8))
                                                                     (boxes within 1e-5 of a level boundary are excluded)
  cell/grid_anchors      anchor_utils.AnchorGenerator                tv rounds the cell anchors to integers (compared after rounding;
                                                                     the grid is compared with tv's cell anchors replaced)
  rpn_proposals          rpn.RegionProposalNetwork.filter_proposals  tv ranks by sigmoid(logit) (monotone; logits kept < 10 so no float ties),
                                                                     removes boxes < 1e-3 (d2: <= 0)
  detections             roi_heads.RoIHeads.postprocess_detections   tv background is class 0 (columns remapped), removes boxes < 1e-2
  backbone               models.resnet101 + FrozenBatchNorm2d        tv strides the 3x3 (v1.5), d2 cfg STRIDE_IN_1X1 (strides moved)
  fpn                    ops.FeaturePyramidNetwork + LastLevelMaxPool none at even sizes
  rpn_head               rpn.RPNHead                                 none
  box_head               faster_rcnn.TwoMLPHead + FastRCNNPredictor  background column first in tv
  mask_head              mask_rcnn.MaskRCNNHeads + MaskRCNNPredictor + roi_heads.maskrcnn_inference   labels 1-based in tv
  roi_pool               (torchvision.ops.roi_align is the op detectron2's ROIAlign itself dispatches to)
  paste_masks            none in torchvision (it pastes with integer boxes + interpolate, a different algorithm): checked against a
                         direct float64 evaluation of the published rule instead (pixel centre -> mask coordinates -> zero-padded
                         bilinear sample -> >= 0.5), pixels within 1e-5 of the threshold excluded
"""
import math

import pytest
import torch

from oracle import maskrcnn as O

tv = pytest.importorskip("torchvision")
from torchvision.models.detection import _utils as det_utils  # noqa: E402
from torchvision.models.detection.anchor_utils import AnchorGenerator  # noqa: E402
from torchvision.models.detection.faster_rcnn import FastRCNNPredictor, TwoMLPHead  # noqa: E402
from torchvision.models.detection.mask_rcnn import MaskRCNNHeads, MaskRCNNPredictor  # noqa: E402
from torchvision.models.detection.roi_heads import RoIHeads, maskrcnn_inference  # noqa: E402
from torchvision.models.detection.rpn import RegionProposalNetwork, RPNHead, concat_box_prediction_layers  # noqa: E402
from torchvision.ops import FeaturePyramidNetwork  # noqa: E402
from torchvision.ops.feature_pyramid_network import LastLevelMaxPool  # noqa: E402
from torchvision.ops.misc import FrozenBatchNorm2d  # noqa: E402
from torchvision.ops.poolers import LevelMapper  # noqa: E402


@pytest.fixture(scope="module")
def weights():
    return O.synth_weights(0)


def _rand_boxes(g, n, W=1088.0, H=800.0):
    x1 = torch.rand(n, generator=g) * W * 0.8
    y1 = torch.rand(n, generator=g) * H * 0.8
    w = torch.rand(n, generator=g) * W * 0.5 + 1.0
    h = torch.rand(n, generator=g) * H * 0.5 + 1.0
    return torch.stack([x1, y1, x1 + w, y1 + h], 1)


def test_apply_deltas_equals_boxcoder():
    g = torch.Generator().manual_seed(0)
    boxes = _rand_boxes(g, 500)
    for weights, k in (((1.0, 1.0, 1.0, 1.0), 1), ((10.0, 10.0, 5.0, 5.0), 9)):
        deltas = torch.randn((500, 4 * k), generator=g) * 3.0
        deltas[::7, 2::4] = 60.0  # far beyond the clamp log(1000/16) * weight
        coder = det_utils.BoxCoder(weights, bbox_xform_clip=math.log(1000.0 / 16))
        want = coder.decode_single(deltas, boxes)
        got = O.apply_deltas(deltas, boxes, weights)
        assert torch.equal(got, want)


def test_assign_levels_equals_levelmapper():
    g = torch.Generator().manual_seed(1)
    boxes = _rand_boxes(g, 4000)
    # add exact canonical sizes (level boundaries: sqrt(area) = 224 * 2^j)
    for j, s in enumerate((56.0, 112.0, 224.0, 448.0, 896.0)):
        boxes[j] = torch.tensor([10.0, 10.0, 10.0 + s, 10.0 + s])
    s = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    frac = torch.log2(s / 224.0)
    away = (frac - frac.round()).abs() > 1e-5   # the two libraries place their epsilon differently (see the module docstring)
    away[:5] = True                              # ... which does not matter AT the boundary: both round it up
    mapper = LevelMapper(2, 5, canonical_scale=224, canonical_level=4)
    want = mapper([boxes])
    got = O.assign_levels(boxes)
    assert torch.equal(got[away], want[away])
    assert got[:5].tolist() == [0, 1, 2, 3, 3]


def test_anchors_equal_anchor_generator():
    sizes = tuple((s,) for s in O.ANCHOR_SIZES)
    ag = AnchorGenerator(sizes=sizes, aspect_ratios=(O.ANCHOR_RATIOS,) * 5)
    for i, s in enumerate(O.ANCHOR_SIZES):
        want = ag.generate_anchors((s,), O.ANCHOR_RATIOS)          # rounded to integers by torchvision
        assert torch.equal(O.cell_anchors(s).round(), want)
        assert float((O.cell_anchors(s) - want).abs().max()) <= 0.5
    # grid order (y, x, anchor) and shifts: give torchvision the oracle's un-rounded cell anchors
    ag.cell_anchors = [O.cell_anchors(s) for s in O.ANCHOR_SIZES]
    grids = [(25, 34), (13, 17), (7, 9), (4, 5), (2, 3)]
    strides = [[torch.tensor(st), torch.tensor(st)] for st in O.FPN_STRIDES]
    want = ag.grid_anchors(grids, strides)
    for (h, w), st, size, wnt in zip(grids, O.FPN_STRIDES, O.ANCHOR_SIZES, want):
        assert torch.equal(O.grid_anchors(h, w, st, size), wnt)
    assert sum(200 * 272 * 3 // (4 ** i) for i in range(3)) + 25 * 34 * 3 + 13 * 17 * 3 == 217413  # the census of the 800 x 1088 input


def _tv_rpn(pre=1000, post=1000, nms=0.7):
    sizes = tuple((s,) for s in O.ANCHOR_SIZES)
    ag = AnchorGenerator(sizes=sizes, aspect_ratios=(O.ANCHOR_RATIOS,) * 5)
    ag.cell_anchors = [O.cell_anchors(s) for s in O.ANCHOR_SIZES]
    rpn = RegionProposalNetwork(ag, RPNHead(256, 3), 0.7, 0.3, 256, 0.5, dict(training=pre, testing=pre),
                                dict(training=post, testing=post), nms, score_thresh=0.0)
    return rpn.eval()


@pytest.mark.parametrize("pre,post", [(1000, 1000), (50, 30)])
def test_rpn_proposals_equal_filter_proposals(pre, post):
    g = torch.Generator().manual_seed(2)
    image_size = (200, 264)   # not a multiple of the coarsest stride: clipping matters
    grids = [(50, 66), (25, 33), (13, 17), (7, 9), (4, 5)]
    head = []
    for h, w in grids:
        head.append((torch.randn((1, 3, h, w), generator=g) * 2.0, torch.randn((1, 12, h, w), generator=g) * 0.5))
    cfg = O.Cfg(pre_nms_topk=pre, post_nms_topk=post)
    boxes, logits = O.rpn_proposals(head, image_size, cfg)

    rpn = _tv_rpn(pre, post)
    strides = [[torch.tensor(st), torch.tensor(st)] for st in O.FPN_STRIDES]
    anchors = [torch.cat(rpn.anchor_generator.grid_anchors(grids, strides))]
    obj = [o for o, _ in head]
    dlt = [d for _, d in head]
    num_anchors_per_level = [o[0].numel() for o in obj]
    objectness, deltas = concat_box_prediction_layers(obj, dlt)
    proposals = rpn.box_coder.decode(deltas.detach(), anchors).view(1, -1, 4)
    tv_boxes, tv_scores = rpn.filter_proposals(proposals, objectness, [image_size], num_anchors_per_level)
    assert boxes.shape[0] == tv_boxes[0].shape[0] > 10
    assert torch.equal(boxes, tv_boxes[0])
    assert torch.allclose(torch.sigmoid(logits), tv_scores[0], atol=1e-7, rtol=0)


def test_detections_equal_postprocess_detections():
    g = torch.Generator().manual_seed(3)
    R, K = 600, 9
    image_size = (800, 1067)
    proposals = _rand_boxes(g, R)
    logits = torch.randn((R, K + 1), generator=g) * 3.0
    deltas = torch.randn((R, 4 * K), generator=g) * 1.5
    for thresh, ndet in ((0.3, 100), (0.05, 20)):
        cfg = O.Cfg(score_thresh=thresh, detections=ndet)
        boxes, scores, classes = O.detections(logits, deltas, proposals, image_size, cfg)
        heads = RoIHeads(None, None, None, 0.5, 0.5, 512, 0.25, cfg.box_weights, thresh, cfg.box_nms, ndet)
        tv_logits = torch.cat([logits[:, -1:], logits[:, :-1]], 1)            # background first
        tv_reg = torch.cat([torch.zeros((R, 4)), deltas], 1)
        b, s, l = heads.postprocess_detections(tv_logits, tv_reg, [proposals], [image_size])
        assert boxes.shape[0] == b[0].shape[0] and (ndet == 20 or boxes.shape[0] > 20)
        assert torch.equal(classes, l[0] - 1)
        assert torch.equal(boxes, b[0])
        assert torch.allclose(scores, s[0], atol=5e-7, rtol=0)   # softmax over permuted columns: summation order


def _tv_resnet101(w):
    net = tv.models.resnet101(weights=None, norm_layer=FrozenBatchNorm2d)
    sd = {}
    p = "backbone.bottom_up."

    def bn(dst, src):
        for k in ("weight", "bias", "running_mean", "running_var"):
            sd[f"{dst}.{k}"] = w[f"{src}.{k}"]

    sd["conv1.weight"] = w[p + "stem.conv1.weight"]
    bn("bn1", p + "stem.conv1.norm")
    for si, n in enumerate(O.STAGE_BLOCKS):
        for bi in range(n):
            src, dst = f"{p}res{si + 2}.{bi}", f"layer{si + 1}.{bi}"
            for j in (1, 2, 3):
                sd[f"{dst}.conv{j}.weight"] = w[f"{src}.conv{j}.weight"]
                bn(f"{dst}.bn{j}", f"{src}.conv{j}.norm")
            if bi == 0:
                sd[f"{dst}.downsample.0.weight"] = w[f"{src}.shortcut.weight"]
                bn(f"{dst}.downsample.1", f"{src}.shortcut.norm")
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("fc.") for k in missing)
    for layer in (net.layer2, net.layer3, net.layer4):   # STRIDE_IN_1X1 (yaml:111): the 1x1 strides, not the 3x3
        layer[0].conv1.stride, layer[0].conv2.stride = (2, 2), (1, 1)
    return net.eval()


def test_backbone_fpn_rpn_head_equal_torchvision_modules(weights):
    w = weights
    g = torch.Generator().manual_seed(4)
    x = torch.randn((1, 3, 128, 192), generator=g) * 50.0
    with torch.no_grad():
        feats = O.backbone(x, w)
        net = _tv_resnet101(w)
        t = net.maxpool(net.relu(net.bn1(net.conv1(x))))
        tv_feats = {}
        for i, layer in enumerate((net.layer1, net.layer2, net.layer3, net.layer4)):
            t = layer(t)
            tv_feats[f"res{i + 2}"] = t
        for k in feats:
            assert feats[k].shape == tv_feats[k].shape
            assert float((feats[k] - tv_feats[k]).abs().max()) <= 2e-5 * float(tv_feats[k].abs().max()), k
        # FPN
        fpn = FeaturePyramidNetwork([256, 512, 1024, 2048], 256, extra_blocks=LastLevelMaxPool()).eval()
        sd = {}
        for i, lvl in enumerate((2, 3, 4, 5)):
            sd[f"inner_blocks.{i}.0.weight"] = w[f"backbone.fpn_lateral{lvl}.weight"]
            sd[f"inner_blocks.{i}.0.bias"] = w[f"backbone.fpn_lateral{lvl}.bias"]
            sd[f"layer_blocks.{i}.0.weight"] = w[f"backbone.fpn_output{lvl}.weight"]
            sd[f"layer_blocks.{i}.0.bias"] = w[f"backbone.fpn_output{lvl}.bias"]
        fpn.load_state_dict(sd)
        pyr = O.fpn(feats, w)
        from collections import OrderedDict
        tv_pyr = fpn(OrderedDict((k, feats[k]) for k in ("res2", "res3", "res4", "res5")))
        for k, tk in zip(("p2", "p3", "p4", "p5", "p6"), ("res2", "res3", "res4", "res5", "pool")):
            assert float((pyr[k] - tv_pyr[tk]).abs().max()) <= 2e-5 * float(pyr[k].abs().max()), k
        # RPN head
        head = RPNHead(256, 3).eval()
        r = "proposal_generator.rpn_head."
        head.load_state_dict({"conv.0.0.weight": w[r + "conv.weight"], "conv.0.0.bias": w[r + "conv.bias"],
                              "cls_logits.weight": w[r + "objectness_logits.weight"], "cls_logits.bias": w[r + "objectness_logits.bias"],
                              "bbox_pred.weight": w[r + "anchor_deltas.weight"], "bbox_pred.bias": w[r + "anchor_deltas.bias"]})
        got = O.rpn_head(pyr, w)
        obj, dlt = head([pyr[f"p{i}"] for i in range(2, 7)])
        for (o, d), to, td in zip(got, obj, dlt):
            assert float((o - to).abs().max()) <= 1e-4 and float((d - td).abs().max()) <= 1e-4


def test_box_and_mask_heads_equal_torchvision_modules(weights):
    w = weights
    g = torch.Generator().manual_seed(5)
    p = "roi_heads."
    with torch.no_grad():
        pooled = torch.randn((40, 256, 7, 7), generator=g)
        mlp = TwoMLPHead(256 * 49, 1024).eval()
        mlp.load_state_dict({"fc6.weight": w[p + "box_head.fc1.weight"], "fc6.bias": w[p + "box_head.fc1.bias"],
                             "fc7.weight": w[p + "box_head.fc2.weight"], "fc7.bias": w[p + "box_head.fc2.bias"]})
        pred = FastRCNNPredictor(1024, 10).eval()
        cw, cb = w[p + "box_predictor.cls_score.weight"], w[p + "box_predictor.cls_score.bias"]
        bw, bb = w[p + "box_predictor.bbox_pred.weight"], w[p + "box_predictor.bbox_pred.bias"]
        pred.load_state_dict({"cls_score.weight": torch.cat([cw[-1:], cw[:-1]]), "cls_score.bias": torch.cat([cb[-1:], cb[:-1]]),
                              "bbox_pred.weight": torch.cat([torch.zeros((4, 1024)), bw]), "bbox_pred.bias": torch.cat([torch.zeros(4), bb])})
        tl, tr = pred(mlp(pooled))
        logits, deltas = O.box_head(pooled, w)
        assert torch.allclose(logits, torch.cat([tl[:, 1:], tl[:, :1]], 1), atol=2e-5, rtol=1e-5)
        assert torch.allclose(deltas, tr[:, 4:], atol=2e-5, rtol=1e-5)

        mp = torch.randn((12, 256, 14, 14), generator=g)
        classes = torch.randint(0, 9, (12,), generator=g)
        heads = MaskRCNNHeads(256, (256, 256, 256, 256), 1).eval()
        heads.load_state_dict({f"{i}.0.{k}": w[f"{p}mask_head.mask_fcn{i + 1}.{k}"] for i in range(4) for k in ("weight", "bias")})
        mpred = MaskRCNNPredictor(256, 256, 10).eval()
        pw, pb = w[p + "mask_head.predictor.weight"], w[p + "mask_head.predictor.bias"]
        mpred.load_state_dict({"conv5_mask.weight": w[p + "mask_head.deconv.weight"], "conv5_mask.bias": w[p + "mask_head.deconv.bias"],
                               "mask_fcn_logits.weight": torch.cat([torch.zeros_like(pw[:1]), pw]),
                               "mask_fcn_logits.bias": torch.cat([torch.zeros(1), pb])})
        tv_probs = maskrcnn_inference(mpred(heads(mp)), [classes + 1])[0][:, 0]
        probs = O.mask_head(mp, classes, w)
        assert probs.shape == (12, 28, 28)
        assert torch.allclose(probs, tv_probs, atol=2e-5, rtol=1e-5)


def test_paste_masks_against_direct_evaluation():
    """paste_masks_in_image (detectron2 layers/mask_ops.py): output pixel (y, x) samples the 28 x 28 mask at the image of its
    CENTRE (x + 0.5, y + 0.5) under the box -> [0, 28) map, bilinearly, zeros outside, and is set when the value is >= 0.5.
    Evaluated here pixel by pixel in float64 without grid_sample."""
    import numpy as np
    g = torch.Generator().manual_seed(11)
    N, S, H, W = 6, 28, 40, 52
    masks = torch.rand((N, S, S), generator=g)
    boxes = torch.tensor([[3.2, 4.7, 30.9, 33.1], [0.0, 0.0, 52.0, 40.0], [-6.5, -3.0, 20.0, 18.5], [40.0, 30.0, 60.0, 47.0],
                          [10.0, 10.0, 11.5, 31.0], [25.3, 2.2, 50.1, 9.9]])
    got = O.paste_masks(masks, boxes, H, W, 0.5).numpy()
    m = masks.double().numpy()
    checked = 0
    for n in range(N):
        x0, y0, x1, y1 = [float(v) for v in boxes[n]]
        for y in range(H):
            v = (y + 0.5 - y0) / (y1 - y0) * S - 0.5          # mask row coordinate of the pixel centre
            for x in range(W):
                u = (x + 0.5 - x0) / (x1 - x0) * S - 0.5
                fu, fv = math.floor(u), math.floor(v)
                val = 0.0
                for dv in (0, 1):
                    for du in (0, 1):
                        r, c = fv + dv, fu + du
                        if 0 <= r < S and 0 <= c < S:
                            val += m[n, r, c] * (1 - abs(v - r)) * (1 - abs(u - c))
                if abs(val - 0.5) > 1e-5:
                    assert bool(got[n, y, x]) == (val >= 0.5), (n, y, x, val)
                    checked += 1
    assert checked > 0.99 * N * H * W
    assert got[1].any() and not got[3, :25].any()
