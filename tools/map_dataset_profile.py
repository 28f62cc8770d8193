"""Timing of the N4 device kernels (pn_map_quantize, pn_map_sample) at the reference's sizes against the HBM roofline:
CUDA events on the launching stream, 20 launches after 3 warm-ups.  The 20-step sequence (258 MB) and the 32-map batch exceed
L2; the single map (51.6 MB) does not, and its line says so."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from peanut_b200 import map_dataset as D

peak = 6549.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


g = torch.Generator(device="cuda").manual_seed(1)
for E in (1, 32):
    full = torch.rand((E, 14, 960, 960), generator=g, device="cuda")
    t = timed(lambda: D.quantize_full_map_device(full))
    b = full.numel() * 5
    print(f"pn_map_quantize {E} x 14 x 960 x 960: {t * 1e6:.1f} us, {b / t / 1e9:.0f} GB/s algorithmic (5 B per cell) = {b / t / 1e9 / peak:.2f} of {peak:.0f} GB/s"
          f"{' (fits L2)' if b < 100e6 else ''} [incl. torch.empty of the output]")
seq = D.DeviceMapSequence(torch.randint(0, 256, (20, 14, 960, 960), generator=g, device="cuda", dtype=torch.uint8))
cells = 960 * 960
for kw, bytes_per_cell, nm in ((dict(), 20 + 56 + 56 + 48, "img + input + target"), (dict(hwc=False), 20 + 56 + 48, "input + target"),
                               (dict(hwc=False, target=False), 14 + 56, "input only")):
    t = timed(lambda: seq.sample(3, **kw))
    b = cells * bytes_per_cell
    print(f"pn_map_sample ({nm}) 14 x 960 x 960 of a 20-step sequence: {t * 1e6:.1f} us, {b / t / 1e9:.0f} GB/s algorithmic = {b / t / 1e9 / peak:.2f} of {peak:.0f} GB/s")
