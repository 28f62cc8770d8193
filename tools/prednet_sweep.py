"""BASELINE.json configs[4]: map-prediction-only throughput, 24 x 240 x 240 partial maps, batch 1..256.

    python tools/prednet_sweep.py [bf16|tf32] [max_batch | b1,b2,...]            (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/prednet_sweep.py bf16 256                                           (N GPUs, one rank per GPU)

Per batch size: device time of the CUDA-graph replay of the whole network (CUDA events, 3 warm-ups, inputs resident, a
256 MiB L2 flush write between timed forwards), max over ranks; maps/s = ranks x batch / time; achieved conv TFLOP/s from the
network's algorithmic FLOPs.  One JSON line per batch size on rank 0.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import prednet as oracle  # synthetic checkpoint only (random-init weights of the reference architecture)
from peanut_b200 import prediction


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    arg = sys.argv[2] if len(sys.argv) > 2 else "256"
    batches = [int(v) for v in arg.split(",")] if "," in arg else [1 << i for i in range(9) if (1 << i) <= int(arg)]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    C, H = 24, 240
    w = oracle.synth_state_dict(C, 6, seed=0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    seg = prediction.init_segmentor(prediction._default_cfg(C, 6), device=f"cuda:{local}", precision=prec, state_dict=w)
    for B in batches:  # a new batch size rebuilds the launch list (and re-tunes the tile shapes) inside the same context
        x = torch.rand((B, C, H, H), device="cuda")
        out = seg.forward_device(x)
        for _ in range(3):
            seg.forward_device(x, out=out)
        iters = 10 if B <= 32 else 5
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        for e0, e1 in evs:
            flush.fill_(1)
            e0.record()
            seg.forward_device(x, out=out)
            e1.record()
        torch.cuda.synchronize()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs) / iters
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if rank == 0:
            fl = seg.flops()
            print(json.dumps({"workload": "configs[4] map-prediction only, 24x240x240", "dtype": prec, "n_gpus": world,
                              "batch_per_gpu": B, "ms_per_forward": round(ms, 4), "maps_per_s": round(world * B / ms * 1e3, 1),
                              "conv_tflops_per_gpu": round(fl / ms / 1e9, 1), "launches": seg.num_launches()}), flush=True)
        del x, out
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
