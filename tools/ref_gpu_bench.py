"""The "reference GPU PyTorch path" of BASELINE.md §3: the oracle restatements of the reference's three stages run in
eager PyTorch (cuDNN / cuBLAS kernels) on one B200, fp32 with PyTorch's default TF32-for-convolutions setting, batch 1,
with the reference's host round-trips (numpy in -> numpy out per stage, CPU expit, per-instance Python loop).
A reported comparator for the >= 10x target, not part of the product and not part of bench.py's timed regions.

Usage: python tools/ref_gpu_bench.py [steps]   (prints one JSON line)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import mapper as OB
from oracle import maskrcnn as OA
from oracle import prednet as OC
from oracle import preproc as OP


def run(steps=20, device="cuda:0", warmup=5):
    """Returns the result dict (bench.py calls this for its `ref_gpu_eager` field)."""
    dev = torch.device(device)
    torch.backends.cudnn.benchmark = True          # nav/pred_model_cfg.py:136 (C); detectron2 leaves it off for (A)
    wa = {k: v.to(dev) for k, v in OA.synth_weights(0).items()}
    wc = OC.synth_state_dict(24, 6, seed=0)
    model_c = OC.build(wc, in_channels=24).to(dev)
    args = OB.default_args(device=dev)
    rgb = OA.synth_rgb(0)
    depth = OP.synth_depth(0)
    delta, maps, poses = OB.synth_state(0, OB.default_args())
    pmap = OC.synth_partial_map(24, 240, 240, seed=1234)
    state = dict(maps=torch.from_numpy(maps)[None].to(dev), poses=torch.from_numpy(poses)[None].to(dev))
    delta_d = torch.from_numpy(delta)[None].to(dev)
    cfg = OA.Cfg(score_thresh=0.95)
    t_stage = dict(A=0.0, B=0.0, C=0.0)

    def step(timed):
        t0 = time.perf_counter()
        with torch.device(dev):                    # the oracle's factory calls land on the GPU
            r = OA.forward(rgb, wa, cfg)
            sem = OA.accumulate(r["masks"], r["scores"], r["classes"], cfg.num_classes, 0.95, 0.985, None, 480, 640)
        sem = sem.cpu().numpy()                    # segmentation.py:62
        t1 = time.perf_counter()
        obs = OP.preprocess_obs(rgb, depth, sem)   # agent_helper.py:175-217 on the host, as in the reference
        with torch.device(dev):
            fp, mp, _, cur = OB.forward(torch.from_numpy(obs)[None].to(dev), delta_d, state["maps"], state["poses"], args)
        state["maps"] = mp
        _ = cur.cpu().numpy()                      # agent_state.py:276
        t2 = time.perf_counter()
        with torch.no_grad():
            logits = model_c(torch.from_numpy(pmap)[None].to(dev))[0].cpu().numpy()
        from scipy.special import expit
        pred = expit(logits)                       # prediction.py:158
        t3 = time.perf_counter()
        if timed:
            t_stage["A"] += t1 - t0
            t_stage["B"] += t2 - t1
            t_stage["C"] += t3 - t2
        return pred

    for _ in range(warmup):
        step(False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(True)
    torch.cuda.synchronize()
    ms = 1000.0 * (time.perf_counter() - t0) / steps
    return {"what": "oracle in eager PyTorch on one B200 (reference GPU path, BASELINE.md section 3)", "frames_per_s": 1000.0 / ms,
            "ms_per_frame": ms, "ms_A_maskrcnn": 1000 * t_stage["A"] / steps, "ms_B_mapper": 1000 * t_stage["B"] / steps,
            "ms_C_prednet": 1000 * t_stage["C"] / steps, "steps": steps, "dtype": "fp32 (cudnn.allow_tf32 default)",
            "torch": torch.__version__, "gpu": torch.cuda.get_device_name(dev)}


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 20)))
