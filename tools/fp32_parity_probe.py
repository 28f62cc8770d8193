"""Measurement behind the tolerances of the strict-parity mode (precision "fp32": every convolution as three tf32 tensor-core
products, no rounding of stored activations) in tests/test_conv_gpu.py, test_prednet_gpu.py and test_maskrcnn_gpu.py:
the CUDA path against float64 convolutions, the fp32 stage-C oracle and the fp32 stage-A oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torchvision.ops import box_iou
from peanut_b200 import _lib, prediction
from peanut_b200 import segmentation as S
from oracle import maskrcnn as O
from oracle import prednet as P
from tests import test_conv_gpu as TC

THR = 0.3
which = sys.argv[1:] or ["conv", "prednet", "maskrcnn"]

if "conv" in which:
    ctx = _lib.Context(0)
    for case in TC.CASES[:17] + TC.SPLIT_CASES[:2]:
        for prec, nm in ((_lib.PN_TF32, "tf32"), (_lib.PN_FP32, "fp32")):
            y, ref = TC._run(ctx, case, prec)
            sc = ref.abs().max().item()
            print(f"conv {case} {nm}: max err / scale {(y - ref).abs().max().item() / sc:.3e}", flush=True)

if "prednet" in which:
    w = P.synth_state_dict(in_channels=14, num_classes=6, seed=0)
    om = P.build(w)
    for hw in ((240, 240),):
        x = P.synth_partial_map(14, hw[0], hw[1], seed=6)
        ref = P.run_inference(om, x)[0]
        rng = np.abs(ref).max()
        for prec in ("tf32", "fp32"):
            seg = prediction.init_segmentor(prediction._default_cfg(14, 6), device="cuda:0", precision=prec, state_dict=w)
            t0 = time.time()
            got = seg.forward_device(torch.from_numpy(x)[None].cuda()).cpu().numpy()[0]
            p = seg.forward_device(torch.from_numpy(x)[None].cuda(), apply_sigmoid=True).cpu().numpy()[0]
            print(f"prednet {hw} {prec}: logits err / range {np.abs(got - ref).max() / rng:.3e}  prob err "
                  f"{np.abs(p - P.get_prediction(om, x)).max():.3e}  argmax agree {(got.argmax(0) == ref.argmax(0)).mean():.6f} "
                  f"({time.time() - t0:.1f} s)", flush=True)
            del seg

if "maskrcnn" in which:
    w = O.synth_weights(0)
    for prec in ("fp32",):
        t0 = time.time()
        e = S.MaskRCNN(w, precision=prec, batch=1, height=480, width=640)
        print(f"maskrcnn {prec} build {time.time() - t0:.1f} s", flush=True)
        for seed in (11, 12, 13):
            frame = O.synth_rgb(seed)
            taps = {}
            ref = O.forward(frame, w, O.Cfg(score_thresh=THR), taps=taps)
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
            neq = int((got != ref_sem).sum())
            eq_bin = float(((got > 0) == (ref_sem > 0)).float().mean())
            r5 = e.read_tap("res5", tuple(taps["feats"]["res5"].shape)).cpu()
            rel5 = float((r5 - taps["feats"]["res5"]).abs().max() / taps["feats"]["res5"].abs().max())
            p2 = e.read_tap("p2", tuple(taps["pyr"]["p2"].shape)).cpu()
            relp2 = float((p2 - taps["pyr"]["p2"]).abs().max() / taps["pyr"]["p2"].abs().max())
            ds = float((sc - rs).abs().max()) if same_order else float("nan")
            db = float((bx - rb).abs().max()) if same_order else float("nan")
            print(f"maskrcnn {prec} seed {seed}: oracle dets {rb.shape[0]} ours {nd} matched {matched} same-order {same_order} "
                  f"max|dscore| {ds:.2e} max|dbox| {db:.2e} sem equal {eq:.6f} ({neq} cells differ) binary-agree {eq_bin:.6f} "
                  f"res5 rel {rel5:.2e} p2 rel {relp2:.2e}", flush=True)
        del e
