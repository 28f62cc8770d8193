"""RGB-D frame -> predicted semantic map through the whole dependent chain on the device, in the three precision modes,
against the same chain composed of the CPU oracles (tests/test_pipeline_gpu.py oracle_chain).  For bf16 the oracle chain is the
fp32 one: this is the distance of the throughput path from the reference's arithmetic, not from its own emulation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import mapper as OB, maskrcnn as OA, prednet as OC, preproc as OP
from peanut_b200.pipeline import PerceptionPipeline
from tests.test_pipeline_gpu import oracle_chain

shape, thr = (14, 96, 96), 0.3
wa, wc = OA.synth_weights(0), OC.synth_state_dict(shape[0], 6, seed=0)
ocm = OC.build(wc)
args = OB.default_args()
refs = {}
for prec in ("fp32", "tf32", "bf16"):
    pipe = PerceptionPipeline(wa, wc, num_envs=1, device="cuda:0", precision=prec, map_shape=shape, mode="dependent")
    pipe.args.sem_pred_prob_thr = thr
    pipe.args.goal_thr = thr
    for seed in (10, 11, 12):
        rgb, depth = OA.synth_rgb(seed), OP.synth_depth(seed)[:, :, 0]
        delta, maps, poses = OB.synth_state(seed, args)
        pmap = torch.zeros((1,) + shape, device="cuda")
        pipe.full_map.zero_()
        sem, fp, new_map, poses_out, pred = pipe.step_device(torch.from_numpy(rgb)[None].cuda(), torch.from_numpy(depth)[None].cuda(),
                                                             torch.from_numpy(delta)[None].cuda(), torch.from_numpy(maps)[None].cuda(),
                                                             torch.from_numpy(poses)[None].cuda().clone(), pmap)
        torch.cuda.synchronize()
        if seed not in refs:
            lmb = tuple(int(v) for v in pipe.lmb[0].tolist())
            refs[seed] = oracle_chain(rgb, depth, delta, maps, poses, wa, ocm, thr, (pipe.nc, pipe.full_w, pipe.full_h), lmb,
                                      (pipe.win_x1, pipe.win_y1, shape[1], shape[2]))
        ref = refs[seed]
        sem_eq = float((sem[0].cpu() == ref["sem"]).float().mean())
        fp_eq = float((fp[0].cpu() == ref["fp"]).float().mean())
        dm = (new_map[0].cpu() - ref["new_map"]).abs()
        dp = np.abs(pred[0].cpu().numpy() - ref["pred"])
        amax = float((pred[0].cpu().numpy().argmax(0) == ref["pred"].argmax(0)).mean())
        print(f"whole chain {prec} frame {seed}: argmax map equal {amax:.6f}  stack equal {sem_eq:.6f}  obstacle map equal {fp_eq:.6f}  local map within 1e-4 on "
              f"{float((dm <= 1e-4).float().mean()):.6f} (max {float(dm.max()):.3e})  predicted map max err {dp.max():.3e} mean {dp.mean():.3e} "
              f"within 2e-3 on {float((dp <= 2e-3).mean()):.6f}", flush=True)
    del pipe
