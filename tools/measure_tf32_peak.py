#!/usr/bin/env python
"""Measured dense TF32 (and bf16) tensor throughput of this GPU, with the recipe MEASURED_PEAKS.json states for bf16:
torch.matmul 8192^3 (2*N^3 FLOP), best of 10 (burst) and back to back for 4 s (sustained).  Used as the roofline
denominator of the fp32-storage / tf32-MMA path (cuBLAS is the yardstick here, not part of the product)."""
import json
import sys
import time

import torch


def measure(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    fl = 2.0 * n ** 3
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    iters = 0
    e0.record()
    while time.perf_counter() - t0 < 4.0:
        for _ in range(20):
            a @ b
        iters += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = fl * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return best, sustained


def main():
    out = {"gpu_name": torch.cuda.get_device_name(0),
           "how": "torch.matmul 8192^3 (2*N^3): best of 10 (burst) and back to back for 4 s (sustained); tf32 = fp32 tensors with allow_tf32"}
    out["tf32_tflops"], out["tf32_tflops_sustained"] = measure(torch.float32, True)
    out["bf16_tflops"], out["bf16_tflops_sustained"] = measure(torch.bfloat16, False)
    json.dump(out, open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout, indent=1)


if __name__ == "__main__":
    main()
