"""Oracle for stage A: Mask-RCNN R101-FPN RGB -> per-category mask stack (``SemanticPredMaskRCNN.get_prediction``).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the network itself lives in detectron2 (third-party, v0.6 - the
only release on the cu111/torch1.10 wheel index pinned by peanut.Dockerfile:15), which is neither vendored
under /root/reference nor installable here.  This file restates detectron2 0.6's published inference
algorithm for the architecture fixed by nav/agent/utils/COCO-InstSeg/mask_rcnn_R_101_cat9.yaml and anchors
on the reference's own call sites:

  * wrapper / accumulation          nav/agent/utils/segmentation.py:30-62
  * config                          nav/agent/utils/COCO-InstSeg/mask_rcnn_R_101_cat9.yaml (line numbers below)
  * DefaultPredictor [ext]          ResizeShortestEdge(800, 1333) via PIL bilinear on uint8 (yaml:28,30), BGR (yaml:26)
  * GeneralizedRCNN.inference [ext] (x - PIXEL_MEAN) / PIXEL_STD (yaml:82-89), pad to a multiple of 32
  * ResNet-101 [ext]                STRIDE_IN_1X1 (yaml:111), FrozenBN eps 1e-5 (yaml:103)
  * FPN [ext]                       sum fuse, no norm, 256 ch, LastLevelMaxPool (yaml:62-70)
  * RPN [ext]                       anchors yaml:45-58, top-k 1000/level, NMS 0.7, post top-k 1000 (yaml:224-256)
  * StandardROIHeads [ext]          ROIAlignV2 7x7 / 14x14, 2 FC 1024, weights (10,10,5,5), NMS 0.5, top 100 (yaml:145-223,312)
  * detector_postprocess / paste_masks_in_image [ext]   threshold 0.5

Every function works on plain tensors; ``weights`` is a dict with detectron2 checkpoint key names
 so a real ``model_final.pth['model']`` drops in.
Third-party helpers used as-is (they are the very ops detectron2 0.6 calls): ``torchvision.ops.roi_align`` (aligned=True),
``torchvision.ops.nms``; ``PIL.Image.resize`` for the input resize.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PIXEL_MEAN = (103.53, 116.28, 123.675)  # yaml:82-85 (BGR)
PIXEL_STD = (1.0, 1.0, 1.0)             # yaml:86-89
STAGE_BLOCKS = (3, 4, 23, 3)            # ResNet-101
ANCHOR_SIZES = (32, 64, 128, 256, 512)  # yaml:52-58
ANCHOR_RATIOS = (0.5, 1.0, 2.0)         # yaml:48-51
FPN_STRIDES = (4, 8, 16, 32, 64)
SCALE_CLAMP = math.log(1000.0 / 16)


class Cfg:
    """The knobs of the yaml that shape the inference path (defaults = the reference's)."""

    def __init__(self, **kw):
        self.min_size = 800          # INPUT.MIN_SIZE_TEST yaml:30
        self.max_size = 1333         # INPUT.MAX_SIZE_TEST yaml:28
        self.pre_nms_topk = 1000     # RPN.PRE_NMS_TOPK_TEST yaml:251
        self.post_nms_topk = 1000    # RPN.POST_NMS_TOPK_TEST yaml:249
        self.rpn_nms = 0.7           # RPN.NMS_THRESH yaml:247
        self.num_classes = 9         # ROI_HEADS.NUM_CLASSES yaml:193
        self.score_thresh = 0.95     # ROI_HEADS.SCORE_THRESH_TEST <- args.sem_pred_prob_thr (segmentation.py:33)
        self.box_nms = 0.5           # ROI_HEADS.NMS_THRESH_TEST yaml:192
        self.detections = 100        # TEST.DETECTIONS_PER_IMAGE yaml:312
        self.box_weights = (10.0, 10.0, 5.0, 5.0)  # yaml:145-149
        self.mask_thresh = 0.5
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


# ------------------------------------------------------------------------------------------------ input
def resized_shape(h, w, cfg):
    """ResizeShortestEdge.get_output_shape [ext]: scale shortest edge to min_size, cap longest at max_size."""
    scale = cfg.min_size * 1.0 / min(h, w)
    if h < w:
        newh, neww = cfg.min_size, scale * w
    else:
        newh, neww = scale * h, cfg.min_size
    if max(newh, neww) > cfg.max_size:
        scale = cfg.max_size * 1.0 / max(newh, neww)
        newh, neww = newh * scale, neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def preprocess(img_rgb, cfg):
    """uint8 RGB [H,W,3] -> (normalised, padded BGR float32 [1,3,Hp,Wp], (newh, neww)).
    segmentation.py:44 flips to BGR; DefaultPredictor resizes with PIL bilinear on uint8."""
    from PIL import Image
    img = np.ascontiguousarray(img_rgb[:, :, ::-1])
    h, w = img.shape[:2]
    newh, neww = resized_shape(h, w, cfg)
    pil = Image.fromarray(img)
    pil = pil.resize((neww, newh), Image.BILINEAR)
    x = torch.as_tensor(np.asarray(pil).astype("float32").transpose(2, 0, 1))
    mean = torch.tensor(PIXEL_MEAN).view(3, 1, 1)
    std = torch.tensor(PIXEL_STD).view(3, 1, 1)
    x = (x - mean) / std
    hp, wp = (newh + 31) // 32 * 32, (neww + 31) // 32 * 32
    out = torch.zeros((1, 3, hp, wp))
    out[0, :, :newh, :neww] = x
    return out, (newh, neww)


# ------------------------------------------------------------------------------------------------ backbone
def _frozen_bn(w, prefix):
    scale = w[prefix + ".weight"] * (w[prefix + ".running_var"] + 1e-5).rsqrt()
    bias = w[prefix + ".bias"] - w[prefix + ".running_mean"] * scale
    return scale, bias


def _conv_bn(x, w, name, stride=1, padding=0, relu=False, emulate=False):
    s, b = _frozen_bn(w, name + ".norm")
    if emulate:  # the CUDA path stores bf16(weight * scale) and adds the shift in the epilogue
        y = F.conv2d(x, _r(w[name + ".weight"] * s[:, None, None, None], emulate), None, stride, padding)
        y = y + b[None, :, None, None]
    else:
        y = F.conv2d(x, w[name + ".weight"], None, stride, padding)
        y = y * s[None, :, None, None] + b[None, :, None, None]
    return F.relu(y) if relu else y


def _r(t, emulate):
    """Storage rounding of the CUDA path, applied wherever it materialises a tensor: True / "bf16" = bf16 (round to nearest
    even); "tf32" = the 10-bit tf32 mantissa, round to nearest with ties away from zero (cvt.rna.tf32.f32), i.e. what the
    fp32-storage path keeps so that the tensor core's operand truncation is exact; False = nothing (the fp32 reference)."""
    if not emulate:
        return t
    if emulate == "tf32":
        u = t.contiguous().view(torch.int32)
        return ((u + 0x1000) & -0x2000).view(torch.float32)
    return t.to(torch.bfloat16).to(torch.float32)


def backbone(x, w, emulate_bf16=False):
    """ResNet-101 bottom-up [ext]; returns {res2..res5}.  emulate_bf16 (True / "bf16" / "tf32", see _r) rounds every
    materialised tensor the way the CUDA path stores it."""
    e = emulate_bf16
    p = "backbone.bottom_up."
    x = _r(x, e)
    x = _r(_conv_bn(x, w, p + "stem.conv1", 2, 3, relu=True, emulate=e), e)
    x = F.max_pool2d(x, 3, 2, 1)
    outs = {}
    for si, nblocks in enumerate(STAGE_BLOCKS):
        stage = f"res{si + 2}"
        for bi in range(nblocks):
            pre = f"{p}{stage}.{bi}"
            stride = 2 if (bi == 0 and si > 0) else 1
            if (pre + ".shortcut.weight") in w:
                identity = _r(_conv_bn(x, w, pre + ".shortcut", stride, 0, emulate=e), e)
            else:
                identity = x
            t = _r(_conv_bn(x, w, pre + ".conv1", stride, 0, relu=True, emulate=e), e)  # STRIDE_IN_1X1
            t = _r(_conv_bn(t, w, pre + ".conv2", 1, 1, relu=True, emulate=e), e)
            t = _conv_bn(t, w, pre + ".conv3", 1, 0, emulate=e)
            x = _r(F.relu(t + identity), e)
        outs[stage] = x
    return outs


def fpn(feats, w, emulate_bf16=False):
    """FPN top-down [ext] -> {p2..p6}."""
    e = emulate_bf16
    out = {}
    prev = None
    for lvl in (5, 4, 3, 2):
        lat = F.conv2d(feats[f"res{lvl}"], _r(w[f"backbone.fpn_lateral{lvl}.weight"], e), w[f"backbone.fpn_lateral{lvl}.bias"])
        if prev is not None:
            lat = lat + F.interpolate(prev, scale_factor=2.0, mode="nearest")
        prev = _r(lat, e)
        out[f"p{lvl}"] = _r(F.conv2d(prev, _r(w[f"backbone.fpn_output{lvl}.weight"], e), w[f"backbone.fpn_output{lvl}.bias"],
                                     padding=1), e)
    out["p6"] = F.max_pool2d(out["p5"], kernel_size=1, stride=2, padding=0)
    return out


# ------------------------------------------------------------------------------------------------ RPN
def rpn_head(pyr, w, emulate_bf16=False):
    """StandardRPNHead [ext]: per level (objectness [1,A,H,W], deltas [1,4A,H,W])."""
    e = emulate_bf16
    p = "proposal_generator.rpn_head."
    out = []
    for lvl in range(2, 7):
        t = _r(F.relu(F.conv2d(pyr[f"p{lvl}"], _r(w[p + "conv.weight"], e), w[p + "conv.bias"], padding=1)), e)
        out.append((F.conv2d(t, _r(w[p + "objectness_logits.weight"], e), w[p + "objectness_logits.bias"]),
                    F.conv2d(t, _r(w[p + "anchor_deltas.weight"], e), w[p + "anchor_deltas.bias"])))
    return out


def cell_anchors(size):
    """DefaultAnchorGenerator.generate_cell_anchors [ext] (float64 maths, stored as float32)."""
    a = []
    area = size ** 2.0
    for r in ANCHOR_RATIOS:
        ww = math.sqrt(area / r)
        hh = r * ww
        a.append([-ww / 2.0, -hh / 2.0, ww / 2.0, hh / 2.0])
    return torch.tensor(a, dtype=torch.float32)


def grid_anchors(H, W, stride, size):
    """[H*W*A, 4], order (y, x, anchor); offset 0 (yaml:47)."""
    sx = torch.arange(0, W * stride, step=stride, dtype=torch.float32)
    sy = torch.arange(0, H * stride, step=stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
    return (shifts.view(-1, 1, 4) + cell_anchors(size).view(1, -1, 4)).reshape(-1, 4)


def apply_deltas(deltas, boxes, weights):
    """Box2BoxTransform.apply_deltas [ext]; deltas [N, 4k], boxes [N, 4]."""
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = torch.clamp(deltas[:, 2::4] / ww, max=SCALE_CLAMP)
    dh = torch.clamp(deltas[:, 3::4] / wh, max=SCALE_CLAMP)
    pcx = dx * widths[:, None] + ctr_x[:, None]
    pcy = dy * heights[:, None] + ctr_y[:, None]
    pw = torch.exp(dw) * widths[:, None]
    ph = torch.exp(dh) * heights[:, None]
    x1, y1, x2, y2 = pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph
    return torch.stack((x1, y1, x2, y2), dim=-1).reshape(deltas.shape)


def _clip(boxes, size):
    h, w = size
    boxes = boxes.clone()
    boxes[:, 0].clamp_(min=0, max=w)
    boxes[:, 1].clamp_(min=0, max=h)
    boxes[:, 2].clamp_(min=0, max=w)
    boxes[:, 3].clamp_(min=0, max=h)
    return boxes


def _sorted_desc(scores):
    """Descending order with ties broken by the lower index (torch.sort's CPU behaviour with stable=True)."""
    return torch.sort(scores, descending=True, stable=True)


def rpn_proposals(head_out, image_size, cfg):
    """RPN.predict_proposals + find_top_rpn_proposals [ext] -> (boxes [R,4], logits [R])."""
    from torchvision.ops import nms
    all_boxes, all_scores, all_lvl = [], [], []
    for li, (obj, dlt) in enumerate(head_out):
        _, A, H, W = obj.shape
        logits = obj.permute(0, 2, 3, 1).reshape(-1)                                   # (y, x, a)
        deltas = dlt.view(1, A, 4, H, W).permute(0, 3, 4, 1, 2).reshape(-1, 4)
        anchors = grid_anchors(H, W, FPN_STRIDES[li], ANCHOR_SIZES[li])
        k = min(logits.numel(), cfg.pre_nms_topk)
        s, idx = _sorted_desc(logits)
        s, idx = s[:k], idx[:k]
        boxes = apply_deltas(deltas[idx], anchors[idx], (1.0, 1.0, 1.0, 1.0))
        all_boxes.append(boxes)
        all_scores.append(s)
        all_lvl.append(torch.full((k,), li, dtype=torch.int64))
    boxes, scores, lvl = torch.cat(all_boxes), torch.cat(all_scores), torch.cat(all_lvl)
    valid = torch.isfinite(boxes).all(1) & torch.isfinite(scores)
    boxes, scores, lvl = boxes[valid], scores[valid], lvl[valid]
    boxes = _clip(boxes, image_size)
    keep = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)  # MIN_SIZE 0 (yaml:91)
    boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
    kept = []
    for li in range(len(head_out)):  # batched_nms == independent NMS per level
        m = torch.nonzero(lvl == li).reshape(-1)
        kept.append(m[nms(boxes[m], scores[m], cfg.rpn_nms)])
    kept = torch.cat(kept)
    kept = kept[_sorted_desc(scores[kept])[1]][:cfg.post_nms_topk]
    return boxes[kept], scores[kept]


# ------------------------------------------------------------------------------------------------ ROI heads
def assign_levels(boxes):
    """ROIPooler.assign_boxes_to_levels [ext]: canonical box 224 at level 4, levels 2..5 -> 0..3."""
    sizes = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    lv = torch.floor(4 + torch.log2(sizes / 224 + 1e-8))
    return torch.clamp(lv, min=2, max=5).to(torch.int64) - 2


def roi_pool(pyr, boxes, out_size):
    """ROIPooler with ROIAlignV2 (aligned=True, sampling_ratio 0) over p2..p5 -> [R, 256, S, S]."""
    from torchvision.ops import roi_align
    R = boxes.shape[0]
    out = torch.zeros((R, pyr["p2"].shape[1], out_size, out_size))
    if R == 0:
        return out
    lv = assign_levels(boxes)
    for l in range(4):
        inds = torch.nonzero(lv == l).reshape(-1)
        if inds.numel() == 0:
            continue
        rois = torch.cat((torch.zeros((inds.numel(), 1)), boxes[inds]), dim=1)
        out[inds] = roi_align(pyr[f"p{l + 2}"], rois, out_size, spatial_scale=1.0 / FPN_STRIDES[l], sampling_ratio=0,
                              aligned=True)
    return out


def box_head(pooled, w, emulate_bf16=False):
    """FastRCNNConvFCHead (2 FC) + FastRCNNOutputLayers [ext] -> (class logits [R, K+1], deltas [R, 4K])."""
    e = emulate_bf16
    p = "roi_heads."
    x = _r(pooled, e).flatten(1)
    x = _r(F.relu(F.linear(x, _r(w[p + "box_head.fc1.weight"], e), w[p + "box_head.fc1.bias"])), e)
    x = _r(F.relu(F.linear(x, _r(w[p + "box_head.fc2.weight"], e), w[p + "box_head.fc2.bias"])), e)
    return (F.linear(x, _r(w[p + "box_predictor.cls_score.weight"], e), w[p + "box_predictor.cls_score.bias"]),
            F.linear(x, _r(w[p + "box_predictor.bbox_pred.weight"], e), w[p + "box_predictor.bbox_pred.bias"]))


def detections(cls_logits, deltas, proposals, image_size, cfg):
    """fast_rcnn_inference_single_image [ext] -> (boxes [N,4], scores [N], classes [N]) sorted by score."""
    from torchvision.ops import nms
    K = cfg.num_classes
    boxes = apply_deltas(deltas, proposals, cfg.box_weights)          # [R, 4K]
    probs = F.softmax(cls_logits, dim=-1)
    valid = torch.isfinite(boxes).all(1) & torch.isfinite(probs).all(1)
    boxes, probs = boxes[valid], probs[valid]
    scores = probs[:, :-1]
    boxes = _clip(boxes.reshape(-1, 4), image_size).view(-1, K, 4)
    mask = scores > cfg.score_thresh
    inds = mask.nonzero()
    boxes, scores = boxes[mask], scores[mask]
    kept = []
    for c in range(K):
        m = torch.nonzero(inds[:, 1] == c).reshape(-1)
        if m.numel():
            kept.append(m[nms(boxes[m], scores[m], cfg.box_nms)])
    kept = torch.cat(kept) if kept else torch.zeros((0,), dtype=torch.int64)
    kept = kept[_sorted_desc(scores[kept])[1]][:cfg.detections]
    return boxes[kept], scores[kept], inds[kept, 1]


def mask_head(pooled, classes, w, emulate_bf16=False):
    """MaskRCNNConvUpsampleHead + mask_rcnn_inference [ext]: [N,256,14,14] -> probabilities [N,28,28]."""
    e = emulate_bf16
    p = "roi_heads.mask_head."
    x = _r(pooled, e)
    for i in range(1, 5):
        x = _r(F.relu(F.conv2d(x, _r(w[p + f"mask_fcn{i}.weight"], e), w[p + f"mask_fcn{i}.bias"], padding=1)), e)
    x = _r(F.relu(F.conv_transpose2d(x, _r(w[p + "deconv.weight"], e), w[p + "deconv.bias"], stride=2)), e)
    logits = F.conv2d(x, _r(w[p + "predictor.weight"], e), w[p + "predictor.bias"])
    N = logits.shape[0]
    return logits[torch.arange(N), classes].sigmoid()


def paste_masks(masks, boxes, out_h, out_w, threshold=0.5):
    """paste_masks_in_image / _do_paste_mask [ext]: [N,28,28] probabilities -> bool [N,out_h,out_w]."""
    N = masks.shape[0]
    if N == 0:
        return torch.zeros((0, out_h, out_w), dtype=torch.bool)
    x0, y0, x1, y1 = torch.split(boxes, 1, dim=1)
    img_y = torch.arange(0, out_h, dtype=torch.float32) + 0.5
    img_x = torch.arange(0, out_w, dtype=torch.float32) + 0.5
    img_y = (img_y - y0) / (y1 - y0) * 2 - 1
    img_x = (img_x - x0) / (x1 - x0) * 2 - 1
    gx = img_x[:, None, :].expand(N, img_y.size(1), img_x.size(1))
    gy = img_y[:, :, None].expand(N, img_y.size(1), img_x.size(1))
    grid = torch.stack([gx, gy], dim=3)
    out = F.grid_sample(masks[:, None], grid, align_corners=False)
    return out[:, 0] >= threshold


def postprocess(boxes, scores, classes, mask_probs, image_size, out_h, out_w, cfg):
    """detector_postprocess [ext]: rescale boxes to the original frame, clip, drop empty, paste masks."""
    sx, sy = out_w / image_size[1], out_h / image_size[0]
    b = boxes.clone()
    b[:, 0::2] *= sx
    b[:, 1::2] *= sy
    b = _clip(b, (out_h, out_w))
    ne = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    b, scores, classes, mask_probs = b[ne], scores[ne], classes[ne], mask_probs[ne]
    return b, scores, classes, paste_masks(mask_probs, b, out_h, out_w, cfg.mask_thresh)


def accumulate(masks, scores, classes, n_cats, sem_pred_prob_thr, goal_thr, goal_cat, H, W):
    """segmentation.py:47-62 -> float32 [H, W, n_cats + 1]."""
    out = torch.zeros((H, W, n_cats + 1))
    for j in range(classes.numel()):
        idx = int(classes[j])
        if idx in range(n_cats):
            if scores[j] < sem_pred_prob_thr:
                continue
            if goal_cat is not None and idx == goal_cat and scores[j] < goal_thr:
                continue
            out[:, :, idx] += masks[j] * 1.
    return out


# ------------------------------------------------------------------------------------------------ end to end
def forward(img_rgb, w, cfg=None, emulate_bf16=False, taps=None):
    """uint8 RGB [H,W,3] -> dict(boxes, scores, classes, masks(bool [N,H,W])).  ``taps`` (dict) receives every
    intermediate tensor for stage-wise parity tests."""
    cfg = cfg or Cfg()
    with torch.no_grad():
        x, image_size = preprocess(img_rgb, cfg)
        feats = backbone(x, w, emulate_bf16)
        pyr = fpn(feats, w, emulate_bf16)
        head = rpn_head(pyr, w, emulate_bf16)
        prop, prop_logits = rpn_proposals(head, image_size, cfg)
        pooled = roi_pool(pyr, prop, 7)
        cls_logits, deltas = box_head(pooled, w, emulate_bf16)
        boxes, scores, classes = detections(cls_logits, deltas, prop, image_size, cfg)
        mpooled = roi_pool(pyr, boxes, 14)
        mprobs = mask_head(mpooled, classes, w, emulate_bf16)
        H, W = img_rgb.shape[:2]
        b, s, c, masks = postprocess(boxes, scores, classes, mprobs, image_size, H, W, cfg)
    if taps is not None:
        taps.update(input=x, image_size=image_size, feats=feats, pyr=pyr, rpn=head, proposals=prop,
                    proposal_logits=prop_logits, pooled=pooled, cls_logits=cls_logits, deltas=deltas, det_boxes=boxes,
                    det_scores=scores, det_classes=classes, mask_pooled=mpooled, mask_probs=mprobs)
    return dict(boxes=b, scores=s, classes=c, masks=masks)


def get_prediction(img_rgb, w, cfg=None, sem_pred_prob_thr=0.95, goal_thr=0.985, goal_cat=None, emulate_bf16=False):
    """SemanticPredMaskRCNN.get_prediction (segmentation.py:41-62) -> (float32 [H,W,n_cats+1], BGR image)."""
    cfg = cfg or Cfg(score_thresh=sem_pred_prob_thr)
    r = forward(img_rgb, w, cfg, emulate_bf16)
    H, W = img_rgb.shape[:2]
    sem = accumulate(r["masks"], r["scores"], r["classes"], cfg.num_classes, sem_pred_prob_thr, goal_thr, goal_cat, H, W)
    return sem.numpy(), img_rgb[:, :, ::-1]


# ------------------------------------------------------------------------------------------------ synthetic data
def synth_weights(seed=0, num_classes=9, head_gain=1.0):
    """Seeded detectron2-style checkpoint dict.  Convs Kaiming-normal (fan_out); FrozenBN
    gamma~U(0.5,1.5) (x0.25 on conv3 so 33 residual blocks stay O(1)), beta, mean ~N(0,0.1), var~U(0.5,1.5);
    FPN/RPN/head convs c2_msra-like with small random biases; predictors scaled so that scores spread out
    (with detectron2's N(0,0.01) init every class would score 0.1 and nothing could be told apart)."""
    g = torch.Generator().manual_seed(seed)
    w = {}

    def conv(name, cout, cin, k, bias=False, gain=2.0):
        fan_out = cout * k * k
        w[name + ".weight"] = torch.randn((cout, cin, k, k), generator=g) * (gain / fan_out) ** 0.5
        if bias:
            w[name + ".bias"] = torch.randn((cout,), generator=g) * 0.05

    def bn(name, c, gamma_scale=1.0):
        w[name + ".weight"] = (torch.rand((c,), generator=g) + 0.5) * gamma_scale
        w[name + ".bias"] = torch.randn((c,), generator=g) * 0.1
        w[name + ".running_mean"] = torch.randn((c,), generator=g) * 0.1
        w[name + ".running_var"] = torch.rand((c,), generator=g) + 0.5

    p = "backbone.bottom_up."
    conv(p + "stem.conv1", 64, 3, 7)
    bn(p + "stem.conv1.norm", 64)
    # the stem sees inputs of magnitude ~100 (mean-subtracted pixels, std 1): scale so activations are O(1)
    w[p + "stem.conv1.weight"] *= 1.0 / 40.0
    cin = 64
    for si, nblocks in enumerate(STAGE_BLOCKS):
        mid, cout = 64 << si, 256 << si
        for bi in range(nblocks):
            pre = f"{p}res{si + 2}.{bi}"
            if bi == 0:
                conv(pre + ".shortcut", cout, cin, 1)
                bn(pre + ".shortcut.norm", cout)
            conv(pre + ".conv1", mid, cin, 1)
            bn(pre + ".conv1.norm", mid)
            conv(pre + ".conv2", mid, mid, 3)
            bn(pre + ".conv2.norm", mid)
            conv(pre + ".conv3", cout, mid, 1)
            bn(pre + ".conv3.norm", cout, 0.25)
            cin = cout
    def conv_in(name, cout, cin, k, gain, bias_std=0.05):
        """fan-in scaled init with bias: output variance ~ gain * E[x^2]."""
        w[name + ".weight"] = torch.randn((cout, cin, k, k), generator=g) * (gain / (cin * k * k)) ** 0.5
        w[name + ".bias"] = torch.randn((cout,), generator=g) * bias_std

    for lvl, c in zip((2, 3, 4, 5), (256, 512, 1024, 2048)):
        conv_in(f"backbone.fpn_lateral{lvl}", 256, c, 1, gain=0.5 if lvl >= 4 else 8.0)
        conv_in(f"backbone.fpn_output{lvl}", 256, 256, 3, gain=1.0)
    r = "proposal_generator.rpn_head."
    conv_in(r + "conv", 256, 256, 3, gain=2.0)
    conv_in(r + "objectness_logits", 3, 256, 1, gain=8.0 * head_gain, bias_std=0.5)
    conv_in(r + "anchor_deltas", 12, 256, 1, gain=0.15 * head_gain, bias_std=0.1)
    h = "roi_heads."

    def fc(name, cout, cin, gain, bias_std=0.05):
        w[name + ".weight"] = torch.randn((cout, cin), generator=g) * (gain / cin) ** 0.5
        w[name + ".bias"] = torch.randn((cout,), generator=g) * bias_std

    fc(h + "box_head.fc1", 1024, 256 * 7 * 7, 2.0)
    fc(h + "box_head.fc2", 1024, 1024, 2.0)
    fc(h + "box_predictor.cls_score", num_classes + 1, 1024, 6.0 * head_gain, bias_std=0.5)
    fc(h + "box_predictor.bbox_pred", num_classes * 4, 1024, 1.0 * head_gain, bias_std=0.2)
    for i in range(1, 5):
        conv_in(h + f"mask_head.mask_fcn{i}", 256, 256, 3, gain=2.0)
    w[h + "mask_head.deconv.weight"] = torch.randn((256, 256, 2, 2), generator=g) * (2.0 / 256) ** 0.5
    w[h + "mask_head.deconv.bias"] = torch.randn((256,), generator=g) * 0.05
    conv_in(h + "mask_head.predictor", num_classes, 256, 1, gain=4.0 * head_gain, bias_std=0.3)
    # post-ReLU features have a large common positive mean; zero-sum predictor rows keep one class / anchor
    # from winning everywhere, so that top-k, per-class NMS and the class-indexed mask pick see diverse inputs
    for k in (r + "objectness_logits.weight", h + "box_predictor.cls_score.weight", h + "mask_head.predictor.weight"):
        w[k] = w[k] - w[k].mean(dim=tuple(range(1, w[k].dim())), keepdim=True)
    return w


def synth_rgb(seed=0, H=480, W=640):
    """Low-frequency colour noise + jitter: uint8 RGB [H,W,3]."""
    g = torch.Generator().manual_seed(1234 + seed)
    low = torch.rand((1, 3, 15, 20), generator=g) * 255.0
    img = F.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)[0]
    img = img + (torch.rand((3, H, W), generator=g) * 16.0 - 8.0)
    # a few hard-edged rectangles so that box regression / masks see structure
    for _ in range(6):
        h = int(torch.randint(max(2, H // 16), max(3, H * 5 // 12), (1,), generator=g))
        w_ = int(torch.randint(max(2, W // 21), max(3, W * 13 // 32), (1,), generator=g))
        y0, x0 = int(torch.randint(0, H - h, (1,), generator=g)), int(torch.randint(0, W - w_, (1,), generator=g))
        img[:, y0:y0 + h, x0:x0 + w_] = torch.rand((3, 1, 1), generator=g) * 255.0
    return img.clamp(0, 255).round().to(torch.uint8).permute(1, 2, 0).contiguous().numpy()
