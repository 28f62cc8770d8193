"""Per-layer device time of the map-completion network (eager launches + CUDA events)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import prednet as oracle
from peanut_b200 import prediction

def main():
    B, C, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
    w = oracle.synth_state_dict(C, 6, seed=0)
    seg = prediction.init_segmentor(prediction._default_cfg(C, 6), device="cuda:0", precision=prec, state_dict=w)
    x = torch.rand((B, C, H, H), device="cuda")
    seg.forward_device(x); seg.forward_device(x); torch.cuda.synchronize()
    prof = seg.profile(5)
    tot = sum(ms for _, ms, _ in prof)
    print(f"# prednet B={B} C={C} H={H} {prec}: {len(prof)} ops, eager sum {tot:.3f} ms")
    for name, ms, fl in prof:
        print(f"{ms*1000:9.1f} us  {fl/ms/1e9 if ms > 0 else 0:8.1f} TF/s  {name}")
    # graph replay timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = seg.forward_device(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        seg.forward_device(x, out=out)
    e1.record(); torch.cuda.synchronize()
    print(f"# graph replay: {e0.elapsed_time(e1)/10:.3f} ms per forward")

main()
