"""Writes tests/golden/maskrcnn_d2_<case>.npz: outputs of the UNMODIFIED reference stage A - ``SemanticPredMaskRCNN``
(nav/agent/utils/segmentation.py:28-62) on top of detectron2 0.6's ``DefaultPredictor`` - for the oracle's seeded synthetic
weights and frames, next to oracle/maskrcnn.py on the same inputs.

detectron2 is NOT installable in the build container (no wheel, no network), so this script has never run there and the
stage-A oracle stays "parity unpinned" until someone runs it where the reference's own environment exists
(peanut.Dockerfile: torch 1.10 + detectron2 0.6):

    cd <PEANUT checkout> && python <this repo>/tests/golden/make_maskrcnn_golden.py [--device cpu|cuda:0]

It must be started from the reference's root because the wrapper opens its yaml by a relative path (segmentation.py:32).
tests/test_maskrcnn_oracle_cpu.py::test_oracle_matches_detectron2_golden and tests/test_maskrcnn_gpu.py consume the
files when they exist (and say "skipped: no detectron2 golden" otherwise).

What is stored per case (small: sub-sampled activations, full discrete results):
  input_hw, image_size                     network input geometry after ResizeShortestEdge + padding
  res2..res5, p2..p6  [C, ::8, ::8]        sub-sampled activations (fp16) + their full-tensor sums (float64)
  prop_boxes [R,4], prop_logits [R]        RPN proposals in detectron2's order
  det_boxes [N,4], det_scores, det_classes ROI-head detections before detector_postprocess
  sem_u8 [480,640,10]                      the wrapper's output (values are small integers), THE parity target
"""
import argparse
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(os.getcwd(), "nav"))

from oracle import maskrcnn as oracle  # noqa: E402

CASES = [  # (name, frame seed, SCORE_THRESH_TEST = sem_pred_prob_thr, goal_cat, goal_thr)
    ("f11_thr030", 11, 0.3, None, 0.3),
    ("f12_thr030_goal", 12, 0.3, 2, 0.9),
    ("f13_thr095", 13, 0.95, None, 0.985),   # the reference's own thresholds (arguments.py:75-76)
]


def sub(t):
    return t[0, :, ::8, ::8].to(torch.float16).cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    a = ap.parse_args()
    if not os.path.exists("nav/agent/utils/COCO-InstSeg/mask_rcnn_R_101_cat9.yaml"):
        sys.exit("run this from the root of the PEANUT checkout (the wrapper opens its yaml by a relative path)")
    from agent.utils.segmentation import SemanticPredMaskRCNN  # the reference wrapper, unmodified (imports detectron2)

    weights = oracle.synth_weights(0)
    tmp = tempfile.mkdtemp()
    wpath = os.path.join(tmp, "synthetic_model_final.pth")
    torch.save({"model": {k: v.clone() for k, v in weights.items()}}, wpath)   # DetectionCheckpointer: {'model': state_dict}

    for name, seed, thr, goal_cat, goal_thr in CASES:
        frame = oracle.synth_rgb(seed)
        args = types.SimpleNamespace(sem_pred_prob_thr=thr, goal_thr=goal_thr, seg_model_wts=wpath, sem_gpu_id=a.device)
        wrapper = SemanticPredMaskRCNN(args)
        sem, bgr = wrapper.get_prediction(frame, goal_cat=goal_cat)
        assert sem.shape == (480, 640, 10) and np.array_equal(bgr, frame[:, :, ::-1])
        assert (sem == np.round(sem)).all() and sem.max() < 256

        # the same forward once more through detectron2's own modules, for the stage taps
        pred = wrapper.predictor
        model = pred.model
        with torch.no_grad():
            img = np.ascontiguousarray(frame[:, :, ::-1])                      # BGR, as DefaultPredictor receives it
            resized = pred.aug.get_transform(img).apply_image(img)
            t = torch.as_tensor(resized.astype("float32").transpose(2, 0, 1))
            images = model.preprocess_image([{"image": t, "height": 480, "width": 640}])
            bottom = model.backbone.bottom_up(images.tensor)
            feats = model.backbone(images.tensor)
            proposals, _ = model.proposal_generator(images, feats, None)
            instances, _ = model.roi_heads(images, feats, proposals, None)
        inst = instances[0]
        out = dict(seed=seed, thr=thr, goal_cat=-1 if goal_cat is None else goal_cat, goal_thr=goal_thr,
                   input_hw=np.array(images.tensor.shape[2:]), image_size=np.array(images.image_sizes[0]),
                   prop_boxes=proposals[0].proposal_boxes.tensor.cpu().numpy(),
                   prop_logits=proposals[0].objectness_logits.cpu().numpy(),
                   det_boxes=inst.pred_boxes.tensor.cpu().numpy(), det_scores=inst.scores.cpu().numpy(),
                   det_classes=inst.pred_classes.cpu().numpy(), sem_u8=sem.astype(np.uint8),
                   detectron2_version=__import__("detectron2").__version__, torch_version=torch.__version__)
        for k, v in list(bottom.items()) + list(feats.items()):
            out[k] = sub(v)
            out[k + "_sum"] = np.float64(v.double().sum().item())

        # the oracle on the same inputs: report, do not gate (this script is also how a disagreement would be found)
        taps = {}
        cfg = oracle.Cfg(score_thresh=thr)
        ref = oracle.forward(frame, weights, cfg, taps=taps)
        osem = oracle.accumulate(ref["masks"], ref["scores"], ref["classes"], 9, thr, goal_thr, goal_cat, 480, 640).numpy()
        rep = []
        for k in ("res2", "res3", "res4", "res5"):
            rep.append(f"{k} {float((taps['feats'][k] - bottom[k].cpu()).abs().max() / bottom[k].abs().max()):.2e}")
        for k in ("p2", "p3", "p4", "p5", "p6"):
            rep.append(f"{k} {float((taps['pyr'][k] - feats[k].cpu()).abs().max() / feats[k].abs().max()):.2e}")
        same_n = taps["det_boxes"].shape[0] == out["det_boxes"].shape[0]
        print(name, "| rel. diff oracle vs detectron2:", " ".join(rep))
        print(name, "| proposals", taps["proposals"].shape[0], "vs", out["prop_boxes"].shape[0], "| detections",
              taps["det_boxes"].shape[0], "vs", out["det_boxes"].shape[0],
              "| classes equal" if same_n and np.array_equal(taps["det_classes"].numpy(), out["det_classes"]) else "| classes DIFFER",
              "| sem cells differing:", int((osem != sem).sum()), "of", sem.size)
        path = os.path.join(HERE, f"maskrcnn_d2_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
