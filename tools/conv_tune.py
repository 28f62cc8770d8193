"""Times every distinct conv shape of the two networks for each N tile (pn_conv_bench) and prints the best.
Usage: python tools/conv_tune.py [bf16|tf32] [batch]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peanut_b200 import _lib

prec = {"bf16": 0, "tf32": 1}[sys.argv[1] if len(sys.argv) > 1 else "bf16"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
# (name, Cin, H, W, Cout, k, stride, dil, pad, residual)
SHAPES = [
    ("A.res2.conv1", 256, 200, 272, 64, 1, 1, 1, 0, 0), ("A.res2.conv2", 64, 200, 272, 64, 3, 1, 1, 1, 0),
    ("A.res2.conv3", 64, 200, 272, 256, 1, 1, 1, 0, 1), ("A.res3.conv1", 512, 100, 136, 128, 1, 1, 1, 0, 0),
    ("A.res3.conv2", 128, 100, 136, 128, 3, 1, 1, 1, 0), ("A.res3.conv3", 128, 100, 136, 512, 1, 1, 1, 0, 1),
    ("A.res4.conv1", 1024, 50, 68, 256, 1, 1, 1, 0, 0), ("A.res4.conv2", 256, 50, 68, 256, 3, 1, 1, 1, 0),
    ("A.res4.conv3", 256, 50, 68, 1024, 1, 1, 1, 0, 1), ("A.res5.conv1", 2048, 25, 34, 512, 1, 1, 1, 0, 0),
    ("A.res5.conv2", 512, 25, 34, 512, 3, 1, 1, 1, 0), ("A.res5.conv3", 512, 25, 34, 2048, 1, 1, 1, 0, 1),
    ("A.fpn_out2/rpn2", 256, 200, 272, 256, 3, 1, 1, 1, 0), ("A.fpn_out3/rpn3", 256, 100, 136, 256, 3, 1, 1, 1, 0),
    ("A.fpn_out4/rpn4", 256, 50, 68, 256, 3, 1, 1, 1, 0), ("A.fpn_lat2", 256, 200, 272, 256, 1, 1, 1, 0, 0),
    ("A.fc1(M=1000)", 12544, 25, 40, 1024, 1, 1, 1, 0, 0), ("A.fc2(M=1000)", 1024, 25, 40, 1024, 1, 1, 1, 0, 0),
    ("A.mask_fcn(100roi)", 256, 140, 140, 256, 3, 1, 1, 1, 0),
    ("C.l1.conv2", 64, 60, 60, 64, 3, 1, 1, 1, 0), ("C.l2.conv2", 128, 30, 30, 128, 3, 1, 1, 1, 0),
    ("C.l3.conv1", 1024, 30, 30, 256, 1, 1, 1, 0, 0), ("C.l3.conv2", 256, 30, 30, 256, 3, 1, 2, 2, 0),
    ("C.l3.conv3", 256, 30, 30, 1024, 1, 1, 1, 0, 1), ("C.l4.conv1", 2048, 30, 30, 512, 1, 1, 1, 0, 0),
    ("C.l4.conv2", 512, 30, 30, 512, 3, 1, 4, 4, 0), ("C.l4.conv3", 512, 30, 30, 2048, 1, 1, 1, 0, 1),
    ("C.psp.bottleneck", 4096, 30, 30, 512, 3, 1, 1, 1, 0),
]
ctx = _lib.Context(0)
ms, bn = ctypes.c_float(), ctypes.c_int()
print(f"# precision={sys.argv[1] if len(sys.argv) > 1 else 'bf16'} batch={B}; us per launch for N tile auto/32/64/128/256; GF")
for name, cin, h, w, cout, k, st, dil, pad, res in SHAPES:
    if name.startswith("A.fc") or name.startswith("A.mask"):
        b = 1 if B == 1 else B
        hh = h * b if name.startswith("A.fc") else h
        bb = 1
    else:
        bb, hh = B, h
    ho = (hh + 2 * pad - dil * (k - 1) - 1) // st + 1
    wo = (w + 2 * pad - dil * (k - 1) - 1) // st + 1
    gf = 2.0 * bb * ho * wo * cout * cin * k * k / 1e9
    row = []
    for force in (0, 32, 64, 128, 256):
        if force and (force > max(32, (cout + 31) // 32 * 32) or ((cout + 31) // 32 * 32) % force):
            row.append("     -")
            continue
        _lib.check(ctx.lib.pn_conv_bench(ctx.handle, prec, bb, cin, hh, w, cout, k, k, st, dil, pad, res, force, 20,
                                         ctypes.byref(ms), ctypes.byref(bn)))
        row.append(f"{ms.value * 1000:6.1f}" + (f"(bn{bn.value})" if force == 0 else ""))
    print(f"{name:20s} M={bb * ho * wo:7d} N={cout:5d} K={cin * k * k:6d} {gf:7.2f} GF  " + "  ".join(row), flush=True)
