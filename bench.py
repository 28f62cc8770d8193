#!/usr/bin/env python
"""Benchmark of PEANUT's per-step perception hot path (BASELINE.json: frames/s, RGB-D -> predicted semantic map).

One "step" = one perception pass over E environments per GPU in the REFERENCE's order: Mask-RCNN on E 640x480 RGB frames
-> mapper glue + Semantic_Mapping on E depth frames -> update_prediction's stamp + window -> the map-completion net on E
partial maps (BASELINE shape 24x240x240).  Synthetic inputs, seeded random weights of the reference architectures (no
checkpoints exist offline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload auto|cfg1|cfg2|cfg3] [--impl reference]

Workloads (BASELINE.json `configs`):
  cfg2 = 32 frames per GPU, bf16 (configs[2]): the headline at every N - the configuration the north star's roofline target
         is quoted on.  Per-GPU work is FIXED as N grows (weak scaling: value(N) / (N * value(1)) is the efficiency the
         driver computes from the lines); under torchrun (N > 1) every timed step also contains the result gather to rank
         0 (one-sided push over NVLink peer memory, pn_gather_*);
  cfg1 = batch 1, fp32 storage / tf32 MMA (configs[1]): rides along at N = 1 as the nested `latency` block;
  cfg3 = 8 envs per GPU, bf16 (configs[3]: 64 envs on 8 GPUs): measured in the same run at every N as the nested `cfg3`
         block (same step, same gather), so that its own weak-scaling series is in the driver's records too.
Every rank runs the same per-GPU workload on its own environments (no data-path collective).

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
    "cfg2": dict(envs=32, precision="bf16", desc="configs[2]: batch=32 frames per GPU, bf16 tensor-core path"),
}
MAP_SHAPES = {"base": (24, 240, 240), "ref": (14, 720, 720)}
METRIC = "frames/sec (RGB-D -> predicted semantic map)"
VERBOSE = False


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line to stdout when the box
    sets NCCL_DEBUG=VERSION), so file descriptor 1 is pointed at stderr for the whole run and the result line goes to a private
    duplicate of the original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def log(*a):
    if VERBOSE:
        print(f"[bench {time.strftime('%H:%M:%S')}]", *a, file=sys.stderr, flush=True)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    p.add_argument("--map", default="base", choices=sorted(MAP_SHAPES))
    p.add_argument("--mode", default="dependent", choices=["dependent", "overlapped"],
                   help="dependent = the reference's chain A -> B -> C (headline); overlapped = C on the stale map, side stream")
    p.add_argument("--envs", type=int, default=0, help="override environments per GPU")
    p.add_argument("--precision", default="", choices=["", "bf16", "tf32"])
    p.add_argument("--micro-batches", type=int, default=0,
                   help="dependent mode: pipeline the chain over this many groups of environments (0 = the workload's default)")
    p.add_argument("--gather", default="peer", choices=["peer", "nccl", "none"], help="result gather inside the step (N > 1)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-profile", action="store_true")
    p.add_argument("--no-latency", action="store_true", help="skip the nested batch-1 tf32 (configs[1]) block")
    p.add_argument("--no-ref-gpu", action="store_true", help="skip the eager-PyTorch-on-GPU comparator (ref_gpu_eager)")
    p.add_argument("--no-scaling-base", action="store_true", help="skip the nested configs[3] block (cfg3)")
    p.add_argument("--verbose", action="store_true")
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


MODE_DESC = {"dependent": "dependent (reference order: Mask-RCNN -> mapper -> stamp + window -> map completion)",
             "overlapped": "overlapped (map completion on the caller's stale map, side stream)"}


def shared_config(desc, envs, map_shape, mode):
    """The workload description both arms print verbatim (the driver compares the two lines' `config`): what is computed, not how
    it was timed - run bookkeeping goes into `run`."""
    return {"workload": desc, "envs_per_gpu": envs, "frame": [480, 640], "map_shape": list(map_shape), "mode": MODE_DESC[mode],
            "l2": "GPU arm: 256 MiB flush write between timed iterations", "timing": "GPU arm: CUDA events per step, max over ranks"}


def resolve_workload(a, world):
    return a.workload if a.workload != "auto" else "cfg2"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 for its workers; the CPU arm is meant to use every host thread (set before torch loads)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["MKL_NUM_THREADS"] = str(threads)
    map_shape = MAP_SHAPES[a.map]
    step = cpu_step_factory(map_shape, threads)
    times = time_cpu(step, min(a.warmup, 1), a.steps, budget_s=150.0)
    ms = 1000.0 * sum(times) / len(times)
    fps = 1000.0 / ms
    wl = WORKLOADS[resolve_workload(a, world)]
    E = a.envs or wl["envs"]
    sample = (f"{len(times)} of {a.steps} steps executed (150 s budget); each step = 1 frame of the workload's {E} per GPU "
              "(the reference is a single-environment, batch-1, single-process loop: its frames/s does not depend on E)")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": shared_config(wl["desc"], E, map_shape, a.mode),
            "run": {"sample_frames_per_step": 1,
                    "note": "CPU arm: rank 0 only, one frame per step (bounded sample of the same workload)"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------- GPU arm
def load_peaks():
    """MEASURED_PEAKS.json (driver-written: HBM, bf16) + profiles/r02_tf32_peak.json (tools/measure_tf32_peak.py, same recipe)."""
    peaks, src = {}, {}
    try:
        peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
        src["bf16"] = "measured bf16 sustained (MEASURED_PEAKS.json)"
    except Exception:
        peaks["bf16_tflops_sustained"] = 1400.0
        src["bf16"] = "fallback 1400 TFLOP/s bf16 (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_tf32_peak.json")))
        peaks["tf32_tflops_sustained"] = t["tf32_tflops_sustained"]
        src["tf32"] = "measured tf32 sustained: torch.matmul 8192^3 back to back for 4 s on this pool's B200 (profiles/r02_tf32_peak.json)"
    except Exception:
        peaks["tf32_tflops_sustained"] = peaks["bf16_tflops_sustained"] / 2.0
        src["tf32"] = "tf32 = half the bf16 sustained peak (profiles/r02_tf32_peak.json absent)"
    return peaks, src


def measure(a, name, rank, local, world, dev, map_shape, steps, with_clocks, gather_kind):
    """Build the pipeline for workload `name` and time it; returns a dict of measurements (identical on every rank after
    the max-over-ranks reduction)."""
    import torch
    import torch.distributed as dist
    from peanut_b200 import parallel
    from peanut_b200.pipeline import MicroBatchedPipeline, PerceptionPipeline

    wl = WORKLOADS[name]
    E = a.envs or wl["envs"]
    precision = a.precision or wl["precision"]
    wa, wc = synth_weights(map_shape[0])
    log(f"{name}: building E={E} {precision} mode={a.mode}")
    mb = a.micro_batches or wl.get("micro_batches", 1)
    if a.mode != "dependent" or mb < 2 or E % mb != 0:
        mb = 1
    if mb > 1:
        pipe = parallel.build_synchronised(lambda: MicroBatchedPipeline(wa, wc, num_envs=E, micro_batches=mb, device=dev,
                                                                        precision=precision, map_shape=map_shape))
    else:
        pipe = parallel.build_synchronised(lambda: PerceptionPipeline(wa, wc, num_envs=E, device=dev, precision=precision,
                                                                      map_shape=map_shape, mode=a.mode))
    engines = pipe.subs if mb > 1 else [pipe]
    host = synth_inputs(E, map_shape, rank)
    pin = {k: v.pin_memory() for k, v in host.items()}
    d = {k: v.to(dev) for k, v in host.items()}
    maps, poses = d["maps"].clone(), d["poses"].clone()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    gather, gather_impl = None, "none"
    if world > 1 and gather_kind != "none":
        if gather_kind == "peer":
            try:  # raises on every rank together if any rank cannot map the root's slab (CUDA IPC not permitted ...)
                gather, gather_impl = parallel.PeerGather(engines[0].seg.ctx, pipe.pred_out), "peer (pn_gather_*: one-sided NVLink push + flags)"
            except RuntimeError as e:
                print(f"[bench] rank {rank}: {e}; falling back to the NCCL gather", file=sys.stderr, flush=True)
                gather = None
        if gather is None:
            gather_impl = "nccl (dist.gather into pre-allocated buffers)"
            nccl_out = [torch.empty_like(pipe.pred_out) for _ in range(world)] if rank == 0 else None

    def do_gather():
        if world == 1 or gather_kind == "none":
            return
        if gather is not None:
            gather.step(pipe.pred_out)
        else:
            dist.gather(pipe.pred_out, nccl_out, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    def step_dev():
        nonlocal maps
        _, _, maps, _, pred = pipe.step_device(d["rgb"], d["depth"], d["delta"], maps, poses, d["pmap"])
        do_gather()
        return pred

    for i in range(max(a.warmup, 3)):
        step_dev()
        log(f"{name}: warm-up step {i} enqueued")
    barrier()
    log(f"{name}: warm-up done")
    sampler = ClockSampler(local)
    if rank == 0 and with_clocks:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall = time.perf_counter()
    for s0, s1 in ev:
        flush.zero_()                      # evict L2 between timed iterations (not timed)
        s0.record()
        step_dev()
        s1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    dev_ms = sum(s0.elapsed_time(s1) for s0, s1 in ev) / steps
    log(f"{name}: device-resident {dev_ms:.3f} ms/step")

    # ---- gather alone (reported, not subtracted)
    gather_ms = None
    if world > 1 and gather_kind != "none":
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        for _ in range(10):
            do_gather()
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1) / 10.0

    # ---- end to end through the public API with host buffers (`e2e`)
    maps_e, poses_e = d["maps"].clone(), d["poses"].clone()
    for _ in range(3):
        _, _, _, maps_e = pipe.step_host(pin["rgb"], pin["depth"], pin["delta"], pin["pmap"], maps_e, poses_e)
        do_gather()
    barrier()
    e2e_t = []
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, _, maps_e = pipe.step_host(pin["rgb"], pin["depth"], pin["delta"], pin["pmap"], maps_e, poses_e)
        if world > 1 and gather_kind != "none":
            do_gather()
            torch.cuda.synchronize()
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    e2e_ms = 1000.0 * sum(e2e_t) / len(e2e_t)
    clocks = sampler.stop() if (rank == 0 and with_clocks) else None
    log(f"{name}: e2e {e2e_ms:.3f} ms/step")

    vals = [dev_ms, e2e_ms] + ([gather_ms] if gather_ms is not None else [])
    vals = parallel.max_over_ranks(vals, device=dev)
    dev_ms, e2e_ms = vals[0], vals[1]
    if gather_ms is not None:
        gather_ms = vals[2]
    gather_status = gather.status() if gather is not None else 0
    if gather is not None and rank == 0:
        res = gather.result()
        assert tuple(res.shape) == (world,) + tuple(pipe.pred_out.shape)
        assert torch.equal(res[0], pipe.pred_out), "gathered slice of the root differs from its own result"
    if gather_status != 0:
        raise RuntimeError(f"result gather timed out (status {gather_status})")

    out = dict(name=name, E=E, precision=precision, micro_batches=mb, dev_ms=dev_ms, e2e_ms=e2e_ms, gather_ms=gather_ms, gather_impl=gather_impl,
               wall_ms=1000.0 * t_wall / steps, clocks=clocks, launches=pipe.launches_per_step() + (2 if gather is not None else 0),
               h2d=pipe.h2d_bytes(pin["rgb"], pin["depth"], pin["delta"], pin["pmap"]), d2h=pipe.d2h_bytes(), desc=wl["desc"])

    # ---- roofline of the dominant kernel (conv_umma_kernel: every conv / FC of both networks), rank 0
    if rank == 0 and not a.no_profile:
        peaks, src = load_peaks()
        prof_a = [x for e in engines for x in e.seg.profile(3)]
        prof_c = [x for e in engines for x in e.pred.profile(3)]
        # mask-head ops are recorded at capacity (100 ROIs / frame) but run only on the live detections
        ndet = sum(int(e.seg.read_tap("det_count", (e.E,), torch.int32).sum().item()) for e in engines)
        live = ndet / float(E * 100)
        conv_ms, conv_fl, all_ms = 0.0, 0.0, 0.0
        for opname, ms, fl in prof_a + prof_c:
            all_ms += ms
            if fl > 0:
                conv_ms += ms
                conv_fl += fl * (live if opname.startswith("roi_heads.mask_head") else 1.0)
        glue_ms = None
        if a.mode == "dependent":   # glue + mapper + stamp + window: launches outside the two networks' launch lists
            glue_ms = 0.0
            for i, e in enumerate(engines):
                sl = slice(i * e.E, (i + 1) * e.E)
                glue_ms += e.time_glue_mapper_window(d["rgb"][sl], d["depth"][sl], d["delta"][sl], d["maps"][sl], d["poses"][sl].clone(),
                                                     d["pmap"][sl])
            all_ms += glue_ms
        # the ResNet-101 bottom-up alone (the north star quotes its 70 % target on the backbone): conv launches only
        bb = [(ms, fl) for opname, ms, fl in prof_a if opname.startswith("backbone.bottom_up.") and fl > 0]
        bb_ms, bb_fl = sum(x[0] for x in bb), sum(x[1] for x in bb)
        # largest launches of the step that are NOT the conv kernel (name, ms), for the reader of the line
        others = {}
        for opname, ms, fl in prof_a + prof_c:
            if fl <= 0:
                others[opname] = others.get(opname, 0.0) + ms
        top_other = sorted(others.items(), key=lambda kv: -kv[1])[:6]
        n_conv = sum(1 for _, _, fl in prof_a + prof_c if fl > 0)
        achieved = conv_fl / (conv_ms / 1000.0) / 1e12
        key = "bf16" if precision == "bf16" else "tf32"
        peak = peaks[key + "_tflops_sustained"]
        traffic, traffic_src = None, None
        try:  # committed ncu measurement of this workload (profiles/)
            tj = json.load(open(os.path.join(ROOT, "profiles", f"r02_conv_traffic_{name}.json")))
            traffic, traffic_src = tj["traffic_bytes_per_launch"], tj.get("source")
        except Exception:
            pass
        out["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                           "traffic": traffic, "traffic_source": traffic_src, "kernel": "conv_umma_kernel",
                           "launches_per_step": n_conv, "flops_per_step": conv_fl, "kernel_ms_per_step": conv_ms,
                           # share of the timed (dependent: serial; overlapped: two-stream) step that is this kernel's
                           # serialised launch time; > 1 would mean overlap hid part of it
                           "share_of_step": conv_ms / dev_ms, "share_of_serialised_launches": conv_ms / all_ms,
                           "step_flops_over_step_time": conv_fl / (dev_ms / 1000.0) / 1e12,
                           "glue_mapper_window_ms_per_step": glue_ms,
                           "backbone": {"what": "ResNet-101 bottom-up conv launches of stage A", "launches": len(bb), "ms_per_step": bb_ms,
                                        "achieved": bb_fl / (bb_ms / 1000.0) / 1e12 if bb_ms else None,
                                        "frac": bb_fl / (bb_ms / 1000.0) / 1e12 / peak if bb_ms else None},
                           "largest_other_launches_ms": {k: round(v, 4) for k, v in top_other},
                           "detections_per_frame": ndet / float(E), "peak_source": src[key],
                           "how": "CUDA events around every launch (eager replay of the recorded launch list after the timed region)"}
    if gather is not None:
        gather.close()
    del pipe
    torch.cuda.empty_cache()
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist
    from peanut_b200 import parallel

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (ours) needs a GPU: peanut_b200 has no CPU fallback")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    rank, local, world = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    map_shape = MAP_SHAPES[a.map]
    name = resolve_workload(a, world)
    steps = a.steps

    m = measure(a, name, rank, local, world, dev, map_shape, steps, True, a.gather)
    latency = None
    if world == 1 and name != "cfg1" and not a.no_latency:
        latency = measure(a, "cfg1", rank, local, world, dev, map_shape, max(steps, 30), False, "none")

    # configs[3] (8 envs per GPU) in the same run at every N: its weak-scaling series = cfg3.value(N) / (N * cfg3.value(1))
    cfg3 = None
    if name != "cfg3" and a.workload == "auto" and not a.no_scaling_base:
        a3 = argparse.Namespace(**vars(a))
        a3.no_profile = True   # the roofline block belongs to the headline workload
        cfg3 = measure(a3, "cfg3", rank, local, world, dev, map_shape, steps, False, a.gather)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames = m["E"] * world
    line = {"metric": METRIC, "value": frames / (m["dev_ms"] / 1000.0), "unit": "frames/s", "n_gpus": world, "steps": steps,
            "warmup": max(a.warmup, 3), "ms_per_step": m["dev_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": m["precision"], "data": "synthetic",
            "config": shared_config(m["desc"], m["E"], map_shape, a.mode),
            "run": {"micro_batches": m["micro_batches"], "wall_ms_per_step_incl_flush": m["wall_ms"],
                    "gather": m["gather_impl"], "gather_in_step": world > 1 and a.gather != "none", "gather_ms": m["gather_ms"]},
            "e2e": {"value": frames / (m["e2e_ms"] / 1000.0), "unit": "frames/s", "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["e2e_ms"],
                    "note": "pinned host rgb/depth/pose-delta/partial-map in; predicted map, pose, egocentric obstacle map out; "
                            "the per-category masks stay on the device (the glue kernel consumes them)"},
            "gpu_launches": m["launches"] * steps, "clocks": m["clocks"]}
    if "roofline" in m:
        line["roofline"] = m["roofline"]
    if latency is not None:
        line["latency"] = {"workload": latency["desc"], "dtype": latency["precision"], "frames_per_s": 1000.0 / latency["dev_ms"],
                           "ms_per_step": latency["dev_ms"], "e2e_frames_per_s": 1000.0 / latency["e2e_ms"],
                           "e2e_ms_per_step": latency["e2e_ms"], "steps": max(steps, 30),
                           "roofline": latency.get("roofline")}

    if cfg3 is not None:
        line["cfg3"] = {"workload": cfg3["desc"], "envs_per_gpu": cfg3["E"], "n_gpus": world,
                        "value": cfg3["E"] * world / (cfg3["dev_ms"] / 1000.0), "unit": "frames/s",
                        "ms_per_step": cfg3["dev_ms"], "e2e_value": cfg3["E"] * world / (cfg3["e2e_ms"] / 1000.0),
                        "gather_in_step": world > 1 and a.gather != "none", "gather_ms": cfg3["gather_ms"],
                        "note": "BASELINE configs[3] measured like the headline (whole-job frames/s, max over ranks); "
                                "weak-scaling efficiency = cfg3.value(N) / (N * cfg3.value(1))"}

    # ---- the reference's GPU path (eager PyTorch restatement with the reference's host round trips), a reported comparator
    if world == 1 and not a.no_ref_gpu:
        try:
            log("ref_gpu_eager")
            from tools import ref_gpu_bench
            r = ref_gpu_bench.run(steps=10, device=str(dev), warmup=3)
            base = latency["e2e_ms"] if latency is not None else None
            line["ref_gpu_eager"] = {"frames_per_s": r["frames_per_s"], "ms_per_frame": r["ms_per_frame"], "batch": 1,
                                     "stages_ms": {"A": r["ms_A_maskrcnn"], "B": r["ms_B_mapper"], "C": r["ms_C_prednet"]},
                                     "what": r["what"], "dtype": r["dtype"],
                                     "ours_batch1_e2e_over_ref": (r["ms_per_frame"] / base) if base else None,
                                     "ours_headline_e2e_over_ref": line["e2e"]["value"] / r["frames_per_s"]}
        except Exception as e:  # a comparator must never take the bench line down
            line["ref_gpu_eager"] = {"error": str(e)[:200]}

    # ---- CPU baseline on the host cores (bounded sample)
    if not a.no_cpu_baseline and world == 1:
        log("cpu baseline")
        threads = os.cpu_count() or 1
        step = cpu_step_factory(map_shape, threads)
        times = time_cpu(step, 1, 6, budget_s=25.0)
        cms = 1000.0 * sum(times) / len(times)
        line["cpu_baseline"] = {"value": 1000.0 / cms, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"{len(times)} frames, batch 1 fp32, oracle (PyTorch CPU) after 1 warm-up"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    VERBOSE = args.verbose or bool(os.environ.get("PN_BENCH_VERBOSE"))
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        dump = os.environ.get("PN_TUNING_DUMP")  # maintenance: write the launch-configuration table this run ended with
        if dump and int(os.environ.get("RANK", "0")) == 0:
            from peanut_b200 import _lib
            with open(dump, "w") as f:
                f.write(_lib.tuning_export())
