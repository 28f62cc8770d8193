#!/usr/bin/env python
"""Benchmark of PEANUT's per-step perception hot path (BASELINE.json: frames/s, RGB-D -> predicted semantic map).

One "step" = one perception pass over E environments per GPU: Mask-RCNN on E 640x480 RGB frames, the mapper glue +
Semantic_Mapping on E depth frames, and the map-completion net on E partial maps (BASELINE shape 24x240x240).
Synthetic inputs, seeded random weights of the reference architectures (no checkpoints exist offline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg3|cfg2] [--impl reference]

Workloads (BASELINE.json `configs`): cfg1 = batch 1, fp32 storage / tf32 tensor-core operands (configs[1], default);
cfg3 = 8 envs per GPU, bf16 (configs[3] per-GPU slice); cfg2 = 32 frames, bf16 (configs[2]).
Under torchrun every rank runs the same per-GPU workload on its own environments (weak scaling, no data-path
collective; results are gathered with one NCCL all_gather outside the timed region's critical path).

`--impl reference` times the reference's CPU implementation of the same step on the host cores: the reference's own
code where it is importable offline (Semantic_Mapping is restated op by op and pinned bit-exact to it), the
plain-PyTorch restatements of the mmseg / detectron2 networks otherwise (oracle/, kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg1": dict(envs=1, precision="tf32", desc="configs[1]: batch=1 full pipeline (Mask-RCNN + mapper + map-prediction net), fp32 storage / tf32 MMA"),
    "cfg3": dict(envs=8, precision="bf16", desc="configs[3] per-GPU slice: 8 envs per GPU, bf16 tensor-core path"),
    "cfg2": dict(envs=32, precision="bf16", desc="configs[2]: batch=32 frames, bf16 tensor-core path"),
}
MAP_SHAPES = {"base": (24, 240, 240), "ref": (14, 720, 720)}
METRIC = "frames/sec (RGB-D -> predicted semantic map)"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="cfg1", choices=sorted(WORKLOADS))
    p.add_argument("--map", default="base", choices=sorted(MAP_SHAPES))
    p.add_argument("--envs", type=int, default=0, help="override environments per GPU")
    p.add_argument("--precision", default="", choices=["", "bf16", "tf32"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-profile", action="store_true")
    return p.parse_args()


# ----------------------------------------------------------------------------------------------- synthetic data
def synth_inputs(E, map_shape, rank):
    import numpy as np
    import torch
    from oracle import maskrcnn as OA
    from oracle import mapper as OB
    from oracle import prednet as OC
    from oracle import preproc as OP
    args = OB.default_args()
    rgb = np.stack([OA.synth_rgb(rank * 1000 + e) for e in range(E)])
    depth = np.stack([OP.synth_depth(rank * 1000 + e)[:, :, 0] for e in range(E)])
    st = [OB.synth_state(rank * 1000 + e, args) for e in range(E)]
    delta = np.stack([s[0] for s in st])
    maps = np.stack([s[1] for s in st])
    poses = np.stack([s[2] for s in st])
    pmap = np.stack([OC.synth_partial_map(*map_shape, seed=1234 + rank * 1000 + e) for e in range(E)])
    t = torch.from_numpy
    return dict(rgb=t(rgb), depth=t(depth), delta=t(delta), maps=t(maps), poses=t(poses), pmap=t(pmap))


def synth_weights(map_channels):
    from oracle import maskrcnn as OA
    from oracle import prednet as OC
    return OA.synth_weights(0), OC.synth_state_dict(map_channels, 6, seed=0)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU (reference) arm
def cpu_step_factory(map_shape, threads):
    """One frame of the reference's path on the host cores (fp32, batch 1, as the reference runs it)."""
    import numpy as np
    import torch
    from oracle import maskrcnn as OA
    from oracle import mapper as OB
    from oracle import prednet as OC
    from oracle import preproc as OP
    torch.set_num_threads(threads)
    wa, wc = synth_weights(map_shape[0])
    model_c = OC.build(wc, in_channels=map_shape[0])
    args = OB.default_args()
    inp = synth_inputs(1, map_shape, 0)
    state = dict(maps=inp["maps"].clone(), poses=inp["poses"].clone())

    def step():
        rgb = inp["rgb"][0].numpy()
        sem, _ = OA.get_prediction(rgb, wa, sem_pred_prob_thr=0.95, goal_thr=0.985)
        obs = OP.preprocess_obs(rgb, inp["depth"][0].numpy()[:, :, None], sem)
        fp, mp, _, cur = OB.forward(torch.from_numpy(obs)[None], inp["delta"], state["maps"], state["poses"], args)
        state["maps"] = mp
        pred = OC.get_prediction(model_c, inp["pmap"][0].numpy())
        return pred, cur

    return step


def time_cpu(step, warmup, steps, budget_s):
    for _ in range(warmup):
        step()
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    return times


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    map_shape = MAP_SHAPES[a.map]
    step = cpu_step_factory(map_shape, threads)
    times = time_cpu(step, min(a.warmup, 1), a.steps, budget_s=150.0)
    ms = 1000.0 * sum(times) / len(times)
    fps = 1000.0 / ms
    wl = WORKLOADS[a.workload]
    sample = f"{len(times)} of {a.steps} steps executed (150 s budget), 1 frame each, batch 1 fp32"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": wl["desc"], "map_shape": list(map_shape), "frame": [480, 640], "envs_per_gpu": 1,
                       "note": "CPU arm runs batch 1 on rank 0 only (the reference is a single-env, single-process loop)"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from peanut_b200 import parallel
    from peanut_b200.pipeline import PerceptionPipeline

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (ours) needs a GPU: peanut_b200 has no CPU fallback")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    rank, local, world = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    wl = WORKLOADS[a.workload]
    E = a.envs or wl["envs"]
    precision = a.precision or wl["precision"]
    map_shape = MAP_SHAPES[a.map]

    wa, wc = synth_weights(map_shape[0])
    pipe = PerceptionPipeline(wa, wc, num_envs=E, device=dev, precision=precision, map_shape=map_shape)
    host = synth_inputs(E, map_shape, rank)
    pin = {k: v.pin_memory() for k, v in host.items()}
    d = {k: v.to(dev) for k, v in host.items()}
    maps, poses = d["maps"].clone(), d["poses"].clone()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    def step_dev():
        nonlocal maps
        _, _, maps, _, pred = pipe.step_device(d["rgb"], d["depth"], d["delta"], maps, poses, d["pmap"])
        return pred

    for _ in range(max(a.warmup, 3)):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t_wall = time.perf_counter()
    for s0, s1 in ev:
        flush.zero_()                      # evict L2 between timed iterations (not timed)
        s0.record()
        step_dev()
        s1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    dev_ms = sum(s0.elapsed_time(s1) for s0, s1 in ev) / a.steps

    # ---- end to end through the public API with host buffers (`e2e`)
    maps_e, poses_e = d["maps"].clone(), d["poses"].clone()
    for _ in range(3):
        _, _, _, maps_e = pipe.step_host(pin["rgb"], pin["depth"], pin["delta"], pin["pmap"], maps_e, poses_e)
    barrier()
    e2e_t = []
    for _ in range(a.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, _, maps_e = pipe.step_host(pin["rgb"], pin["depth"], pin["delta"], pin["pmap"], maps_e, poses_e)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    e2e_ms = 1000.0 * sum(e2e_t) / len(e2e_t)
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks
    dev_ms, e2e_ms = parallel.max_over_ranks([dev_ms, e2e_ms], device=dev)

    # ---- result gather (the only collective on the path: per-env predicted maps to rank 0; outside the timed region)
    gather_ms = None
    if world > 1:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):  # NCCL sets its point-to-point channels up lazily
            parallel.gather_env_results(pipe.pred_out, E * world, dst=0)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(5):
            allp = parallel.gather_env_results(pipe.pred_out, E * world, dst=0)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = parallel.max_over_ranks([g0.elapsed_time(g1) / 5.0], device=dev)[0]
        assert rank != 0 or tuple(allp.shape) == (E * world,) + tuple(pipe.pred_out.shape[1:])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames = E * world
    line = {"metric": METRIC, "value": frames / (dev_ms / 1000.0), "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": precision, "data": "synthetic",
            "config": {"workload": wl["desc"], "envs_per_gpu": E, "frame": [480, 640], "map_shape": list(map_shape),
                       "l2": "256 MiB flush write between timed iterations", "timing": "CUDA events per step, max over ranks",
                       "wall_ms_per_step_incl_flush": 1000.0 * t_wall / a.steps, "gather_ms": gather_ms},
            "e2e": {"value": frames / (e2e_ms / 1000.0), "unit": "frames/s",
                    "h2d_bytes_per_step": pipe.h2d_bytes(pin["rgb"], pin["depth"], pin["delta"], pin["pmap"]),
                    "d2h_bytes_per_step": pipe.d2h_bytes(), "ms_per_step": e2e_ms},
            "gpu_launches": pipe.launches_per_step() * a.steps, "clocks": clocks}

    # ---- roofline of the dominant kernel (conv_umma_kernel: every conv / FC of both networks)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if not a.no_profile:
        prof_a = pipe.seg.profile(3)
        prof_c = pipe.pred.profile(3)
        # mask-head ops are recorded at capacity (100 ROIs / frame) but run only on the live detections
        ndet = int(pipe.seg.read_tap("det_count", (E,), torch.int32).sum().item())
        live = ndet / float(E * 100)
        conv_ms, conv_fl, all_ms = 0.0, 0.0, 0.0
        for name, ms, fl in prof_a + prof_c:
            all_ms += ms
            if fl > 0:
                conv_ms += ms
                conv_fl += fl * (live if name.startswith("roi_heads.mask_head") else 1.0)
        n_conv = sum(1 for _, _, fl in prof_a + prof_c if fl > 0)
        achieved = conv_fl / (conv_ms / 1000.0) / 1e12
        if precision == "bf16":
            peak, which = peaks.get("bf16_tflops_sustained", 1400.0), "measured bf16 sustained (MEASURED_PEAKS.json)"
        else:
            peak, which = peaks.get("bf16_tflops_sustained", 1400.0) / 2.0, "tf32 = half the measured bf16 sustained peak (tf32 not in MEASURED_PEAKS.json)"
        if not peaks:
            which += " [fallback]"
        traffic, traffic_src = None, None
        if a.workload == "cfg1" and E == 1 and a.map == "base":
            try:  # committed ncu measurement of this exact command (profiles/)
                tj = json.load(open(os.path.join(ROOT, "profiles", "r01_conv_traffic_cfg1.json")))
                traffic, traffic_src = tj["traffic_bytes_per_launch"], "profiles/r01_conv_traffic_cfg1.json (ncu dram__bytes_read+write per conv launch, L2 flushed per kernel)"
            except Exception:
                pass
        line["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                            "traffic": traffic, "traffic_source": traffic_src, "kernel": "conv_umma_kernel", "launches_per_step": n_conv,
                            "flops_per_step": conv_fl, "kernel_ms_per_step": conv_ms, "share_of_step": conv_ms / all_ms,
                            "detections_per_frame": ndet / float(E), "peak_source": which,
                            "how": "CUDA events around every launch (eager replay of the recorded launch list after the timed region)"}

    # ---- CPU baseline on the host cores (bounded sample)
    if not a.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        step = cpu_step_factory(map_shape, threads)
        times = time_cpu(step, 1, 6, budget_s=25.0)
        cms = 1000.0 * sum(times) / len(times)
        line["cpu_baseline"] = {"value": 1000.0 / cms, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"{len(times)} frames, batch 1 fp32, oracle (PyTorch CPU) after 1 warm-up"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
