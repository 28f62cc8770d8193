"""Times the small-M conv shapes of both networks for every (N tile, split-K) pair (pn_conv_bench) next to the automatic
choice.  Usage: python tools/conv_split_sweep.py [bf16|tf32] [batch]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peanut_b200 import _lib

pname = sys.argv[1] if len(sys.argv) > 1 else "bf16"
prec = {"bf16": 0, "tf32": 1}[pname]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
# (name, Cin, H, W, Cout, k, stride, dil, pad, residual)
ONLY = os.environ.get("PN_SWEEP_ONLY", "")  # comma-separated name prefixes
SHAPES = [
    ("A.res2.conv1", 256, 200, 272, 64, 1, 1, 1, 0, 0), ("A.res2.conv2", 64, 200, 272, 64, 3, 1, 1, 1, 0),
    ("A.res2.conv3", 64, 200, 272, 256, 1, 1, 1, 0, 1),
    ("A.res3.conv1", 512, 100, 136, 128, 1, 1, 1, 0, 0), ("A.res3.conv2", 128, 100, 136, 128, 3, 1, 1, 1, 0),
    ("A.res3.conv3", 128, 100, 136, 512, 1, 1, 1, 0, 1),
    ("A.res4.conv1", 1024, 50, 68, 256, 1, 1, 1, 0, 0), ("A.res4.conv2", 256, 50, 68, 256, 3, 1, 1, 1, 0),
    ("A.res4.conv3", 256, 50, 68, 1024, 1, 1, 1, 0, 1), ("A.res5.conv1", 2048, 25, 34, 512, 1, 1, 1, 0, 0),
    ("A.res5.conv2", 512, 25, 34, 512, 3, 1, 1, 1, 0), ("A.res5.conv3", 512, 25, 34, 2048, 1, 1, 1, 0, 1),
    ("A.fpn_out2", 256, 200, 272, 256, 3, 1, 1, 1, 0), ("A.fpn_out3", 256, 100, 136, 256, 3, 1, 1, 1, 0),
    ("A.fpn_out4", 256, 50, 68, 256, 3, 1, 1, 1, 0), ("A.fpn_lat2", 256, 200, 272, 256, 1, 1, 1, 0, 0),
    ("A.mask_fcn", 256, 140, 140, 256, 3, 1, 1, 1, 0),
    ("A.fpn_out5", 256, 25, 34, 256, 3, 1, 1, 1, 0), ("A.fpn_lat5", 2048, 25, 34, 256, 1, 1, 1, 0, 0),
    ("A.fc1(M=1000)", 12544, 25, 40, 1024, 1, 1, 1, 0, 0), ("A.fc2(M=1000)", 1024, 25, 40, 1024, 1, 1, 1, 0, 0),
    ("C.l2.conv2", 128, 30, 30, 128, 3, 1, 1, 1, 0),
    ("C.l3.conv1", 1024, 30, 30, 256, 1, 1, 1, 0, 0), ("C.l3.conv2", 256, 30, 30, 256, 3, 1, 2, 2, 0),
    ("C.l3.conv3", 256, 30, 30, 1024, 1, 1, 1, 0, 1), ("C.l4.conv1", 2048, 30, 30, 512, 1, 1, 1, 0, 0),
    ("C.l4.conv2", 512, 30, 30, 512, 3, 1, 4, 4, 0), ("C.l4.conv3", 512, 30, 30, 2048, 1, 1, 1, 0, 1),
    ("C.psp.bottleneck", 4096, 30, 30, 512, 3, 1, 1, 1, 0),
]
ctx = _lib.Context(0)
ms, bn = ctypes.c_float(), ctypes.c_int()
combos = [(b, s) for b in (64, 128, 256) for s in (1, 2, 4, 8)] + [(b, s) for b in (128, 256) for s in (-1, -2, -4)]  # s < 0: CTA pairs, -s splits
print(f"# precision={pname} batch={B}; us per launch (back-to-back launches, L2-hot): auto | " +
       " ".join(f"{b}/{'P' + str(-s) if s < 0 else s}" for b, s in combos))
for name, cin, h, w, cout, k, st, dil, pad, res in SHAPES:
    if ONLY and not any(name.startswith(o) for o in ONLY.split(",")):
        continue
    bb, hh = (1, h * B) if name.startswith("A.fc") else (B, h)
    ho = (hh + 2 * pad - dil * (k - 1) - 1) // st + 1
    wo = (w + 2 * pad - dil * (k - 1) - 1) // st + 1
    gf = 2.0 * bb * ho * wo * cout * cin * k * k / 1e9
    row = []
    for force in [(0, 0)] + combos:
        b, s = force
        if b and (b > cout or cout % b):
            row.append("   -")
            continue
        force = (b | 0x2000 | (-s << 16)) if s < 0 else (b | 0x4000 | (s << 16)) if b else 0   # pairs / single CTAs with s splits / auto
        rc = ctx.lib.pn_conv_bench(ctx.handle, prec, bb, cin, hh, w, cout, k, k, st, dil, pad, res, force,
                                   20, ctypes.byref(ms), ctypes.byref(bn))
        if rc != 0:
            row.append("   x")
            continue
        row.append(f"{ms.value * 1000:5.1f}" + (f"({bn.value % 1000}/{'P' if bn.value >= 100000 else ''}{(bn.value // 1000) % 100})" if b == 0 else ""))
    print(f"{name:18s} M={bb * ho * wo:6d} N={cout:5d} K={cin * k * k:6d} {gf:6.2f}GF " + " ".join(row), flush=True)
