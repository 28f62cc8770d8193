"""Per-layer roofline of the Mask-RCNN backbone from a per-op profile (tools/mrcnn_profile.py output).

    python tools/layer_roofline.py profiles/r01_ops_maskrcnn_b8_bf16.txt 8 bf16

For every bottom-up ResNet-101 convolution: algorithmic FLOPs and HBM bytes (input read once, output written once,
residual read once, weights once), arithmetic intensity, the roofline bound max(FLOPs / tensor peak, bytes / HBM peak)
with the peaks of MEASURED_PEAKS.json, and the measured time as a fraction of that bound.  Shapes follow
mask_rcnn_R_101_cat9.yaml at the 800 x 1088 network input (480 x 640 frame after ResizeShortestEdge + padding).
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, B, prec = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    es = 2 if prec == "bf16" else 4
    peak_tf = peaks["bf16_tflops_sustained"] * (1.0 if prec == "bf16" else 0.5)
    hbm = peaks["hbm_gbs"]
    rows = []
    for line in open(path):
        m = re.match(r"\s*([\d.]+) us\s+([\d.]+) TF/s\s+backbone\.bottom_up\.res(\d)\.(\d+)\.(\w+)", line)
        if not m:
            continue
        us, tf, S, blk, kind = float(m.group(1)), float(m.group(2)), int(m.group(3)), int(m.group(4)), m.group(5)
        planes = 64 << (S - 2)
        out_c = 4 * planes
        prev_c = 64 if S == 2 else 2 * planes  # channels entering block 0 of the stage
        M = B * (200 >> (S - 2)) * (272 >> (S - 2))
        if kind == "conv1":
            cin = prev_c if blk == 0 else out_c
            K, N, in_b, out_b, res_b = cin, planes, M * cin, M * planes, 0
        elif kind == "conv2":
            K, N, in_b, out_b, res_b = 9 * planes, planes, M * planes, M * planes, 0
        elif kind == "conv3":
            K, N, in_b, out_b, res_b = planes, out_c, M * planes, M * out_c, M * out_c
        else:  # shortcut
            K, N, in_b, out_b, res_b = prev_c, out_c, M * prev_c, M * out_c, 0
        flops = 2.0 * M * K * N
        bytes_ = (in_b + out_b + res_b + K * N) * es
        t_bound = max(flops / (peak_tf * 1e12), bytes_ / (hbm * 1e9)) * 1e6
        rows.append((f"res{S}.{kind}", us, flops, bytes_, t_bound, flops / (peak_tf * 1e12) * 1e6 >= bytes_ / (hbm * 1e9) * 1e6))
    print(f"# {os.path.basename(path)}: batch {B}, {prec}; peaks: tensor {peak_tf:.0f} TFLOP/s, HBM {hbm:.0f} GB/s "
          f"(ridge {peak_tf * 1e3 / hbm:.0f} FLOP/B)")
    print("# layer class      n   time us   TFLOP/s   GB/s   FLOP/B  bound    bound us   time/bound")
    agg = {}
    for name, us, fl, by, tb, tensor in rows:
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0, tensor])
        a[0] += 1; a[1] += us; a[2] += fl; a[3] += by; a[4] += tb
    tot = [0.0, 0.0, 0.0, 0.0]
    for name, (n, us, fl, by, tb, tensor) in agg.items():
        print(f"{name:16s} {n:3d} {us:9.1f} {fl / us / 1e6:9.1f} {by / us / 1e3:6.0f} {fl / by:8.0f}  {'tensor' if tensor else 'hbm':6s} {tb:9.1f} {tb / us:9.2f}")
        tot[0] += us; tot[1] += fl; tot[2] += by; tot[3] += tb
    print(f"{'backbone res2-5':16s}     {tot[0]:9.1f} {tot[1] / tot[0] / 1e6:9.1f} {tot[2] / tot[0] / 1e3:6.0f} {tot[1] / tot[2]:8.0f}  {'':6s} {tot[3]:9.1f} {tot[3] / tot[0]:9.2f}")
    print(f"# tensor-only bound for the same layers: {tot[1] / (peak_tf * 1e12) * 1e6:.1f} us "
          f"({tot[1] / (peak_tf * 1e12) * 1e6 / tot[0]:.2f} of the measured time); roofline bound with HBM: {tot[3]:.1f} us "
          f"({tot[3] / tot[0]:.2f})")


if __name__ == "__main__":
    main()
