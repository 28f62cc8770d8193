"""Per-launch device time of the Mask-RCNN network (eager launches + CUDA events) and graph-replay time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import maskrcnn as O
from peanut_b200 import segmentation as S

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.95
    w = O.synth_weights(0)
    e = S.MaskRCNN(w, precision=prec, batch=B)
    rgb = torch.from_numpy(np.stack([O.synth_rgb(i) for i in range(B)])).cuda()
    out = e.forward_device(rgb, score_thresh=thr, sem_pred_prob_thr=thr)
    out = e.forward_device(rgb, score_thresh=thr, sem_pred_prob_thr=thr, out=out)
    torch.cuda.synchronize()
    nd = e.read_tap("det_count", (B,), torch.int32).cpu().tolist()
    prof = e.profile(3)
    tot = sum(ms for _, ms, _ in prof)
    fl = sum(f for _, _, f in prof)
    print(f"# maskrcnn B={B} {prec}: {len(prof)} ops, eager sum {tot:.3f} ms, capacity FLOPs {fl/1e9:.1f} GF, detections {nd}")
    agg = {}
    for name, ms, f in prof:
        print(f"{ms*1000:9.1f} us  {f/ms/1e9 if ms > 0 else 0:8.1f} TF/s  {name}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        e.forward_device(rgb, score_thresh=thr, sem_pred_prob_thr=thr, out=out)
    e1.record(); torch.cuda.synchronize()
    print(f"# graph replay: {e0.elapsed_time(e1)/10:.3f} ms per forward ({B} frames)")

main()
