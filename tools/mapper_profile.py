"""Device time of stage B (glue + Semantic_Mapping) for E environments: CUDA events around the stage, L2-warm.
Usage: python tools/mapper_profile.py [E] [scene]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import mapper as OB
from peanut_b200 import mapping

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1
scene = sys.argv[2] if len(sys.argv) > 2 else "room"
args = OB.default_args()
args.device = torch.device("cuda:0")
obs = torch.from_numpy(np.stack([OB.synth_obs(e, args, scene, 0.1) for e in range(E)])).cuda()
st = [OB.synth_state(e, args) for e in range(E)]
delta = torch.from_numpy(np.stack([s[0] for s in st])).cuda()
maps = torch.from_numpy(np.stack([s[1] for s in st])).cuda()
poses = torch.from_numpy(np.stack([s[2] for s in st])).cuda()
mod = mapping.Semantic_Mapping(args, num_envs=E)
for _ in range(3):
    mod.forward_batch(obs, delta, maps, poses.clone())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
p = poses.clone()
e0.record()
for _ in range(20):
    mod.forward_batch(obs, delta, maps, p)
e1.record(); torch.cuda.synchronize()
print(f"mapper E={E} scene={scene}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us per forward (7 kernels)")
