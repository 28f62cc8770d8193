"""Measurement behind the end-to-end stage-A tolerances in tests/test_maskrcnn_gpu.py: CUDA Mask-RCNN (tf32 / bf16) against
the oracle with the SAME storage rounding emulated (oracle/maskrcnn.py `_r`), at the reference geometry (480 x 640)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torchvision.ops import box_iou
from oracle import maskrcnn as O
from peanut_b200 import segmentation as S

THR = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
w = O.synth_weights(0)
for prec, emu in (("tf32", "tf32"), ("tf32", False), ("bf16", "bf16")):
    e = S.MaskRCNN(w, precision=prec, batch=1, height=480, width=640)
    for seed in (11, 12, 13):
        frame = O.synth_rgb(seed)
        taps = {}
        ref = O.forward(frame, w, O.Cfg(score_thresh=THR), emulate_bf16=emu, taps=taps)
        sem = e.forward_device(torch.from_numpy(frame)[None].cuda(), score_thresh=THR, sem_pred_prob_thr=THR, goal_thr=THR)
        torch.cuda.synchronize()
        nd = int(e.read_tap("det_count", (1,), torch.int32).item())
        bx = e.read_tap("det_boxes", (100, 4)).cpu()[:nd]
        cl = e.read_tap("det_classes", (100,), torch.int32).cpu()[:nd].long()
        sc = e.read_tap("det_scores", (100,)).cpu()[:nd]
        rb, rc, rs = taps["det_boxes"], taps["det_classes"], taps["det_scores"]
        iou = box_iou(rb, bx) if nd and rb.shape[0] else torch.zeros((rb.shape[0], nd))
        iou[rc[:, None] != cl[None, :]] = 0
        matched = int((iou.max(1).values >= 0.9).sum()) if nd else 0
        same_order = nd == rb.shape[0] and bool((rc == cl).all())
        ref_sem = O.accumulate(ref["masks"], ref["scores"], ref["classes"], 9, THR, THR, None, 480, 640)
        got = sem.cpu()[0]
        eq = float((got == ref_sem).float().mean())
        eq_bin = float(((got > 0) == (ref_sem > 0)).float().mean())
        r5 = e.read_tap("res5", tuple(taps["feats"]["res5"].shape)).cpu()
        rel5 = float((r5 - taps["feats"]["res5"]).abs().max() / taps["feats"]["res5"].abs().max())
        p2 = e.read_tap("p2", tuple(taps["pyr"]["p2"].shape)).cpu()
        relp2 = float((p2 - taps["pyr"]["p2"]).abs().max() / taps["pyr"]["p2"].abs().max())
        ds = float((sc - rs).abs().max()) if same_order else float("nan")
        print(f"{prec} vs oracle(emulate={emu}) seed {seed}: oracle dets {rb.shape[0]} ours {nd} matched {matched} same-order {same_order} "
              f"max|dscore| {ds:.2e} sem equal {eq:.6f} binary-agree {eq_bin:.6f} res5 rel {rel5:.2e} p2 rel {relp2:.2e}", flush=True)
    del e
