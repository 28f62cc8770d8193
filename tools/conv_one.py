"""Runs one named conv shape of the two networks through pn_conv_bench (no torch import: quick to start under ncu).
Usage: python tools/conv_one.py <shape name> [bf16|tf32] [batch] [iters] [force_bn|(splits<<16)]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peanut_b200 import _lib

# name: (Cin, H, W, Cout, k, stride, dil, pad, residual)
SHAPES = {
    "A.stem": (3, 800, 1088, 64, 7, 2, 1, 3, 0),
    "A.res2.conv1": (256, 200, 272, 64, 1, 1, 1, 0, 0), "A.res2.conv2": (64, 200, 272, 64, 3, 1, 1, 1, 0),
    "A.res2.conv3": (64, 200, 272, 256, 1, 1, 1, 0, 1), "A.res3.conv1": (512, 100, 136, 128, 1, 1, 1, 0, 0),
    "A.res3.conv2": (128, 100, 136, 128, 3, 1, 1, 1, 0), "A.res3.conv3": (128, 100, 136, 512, 1, 1, 1, 0, 1),
    "A.res4.conv1": (1024, 50, 68, 256, 1, 1, 1, 0, 0), "A.res4.conv2": (256, 50, 68, 256, 3, 1, 1, 1, 0),
    "A.res4.conv3": (256, 50, 68, 1024, 1, 1, 1, 0, 1), "A.res5.conv1": (2048, 25, 34, 512, 1, 1, 1, 0, 0),
    "A.res5.conv2": (512, 25, 34, 512, 3, 1, 1, 1, 0), "A.res5.conv3": (512, 25, 34, 2048, 1, 1, 1, 0, 1),
    "A.fpn_out2": (256, 200, 272, 256, 3, 1, 1, 1, 0), "A.fpn_out3": (256, 100, 136, 256, 3, 1, 1, 1, 0),
    "A.fpn_lat2": (256, 200, 272, 256, 1, 1, 1, 0, 0),
    "A.mask_fcn": (256, 140, 140, 256, 3, 1, 1, 1, 0),
    "C.l3.conv2": (256, 30, 30, 256, 3, 1, 2, 2, 0), "C.l4.conv2": (512, 30, 30, 512, 3, 1, 4, 4, 0),
    "C.psp.bottleneck": (4096, 30, 30, 512, 3, 1, 1, 1, 0),
}

def main():
    name = sys.argv[1]
    pname = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    force = int(sys.argv[5], 0) if len(sys.argv) > 5 else 0
    cin, h, w, cout, k, st, dil, pad, res = SHAPES[name]
    ctx = _lib.Context(0)
    ms, bn = ctypes.c_float(), ctypes.c_int()
    _lib.check(ctx.lib.pn_conv_bench(ctx.handle, {"bf16": 0, "tf32": 1}[pname], B, cin, h, w, cout, k, k, st, dil, pad, res,
                                     force, iters, ctypes.byref(ms), ctypes.byref(bn)))
    ho = (h + 2 * pad - dil * (k - 1) - 1) // st + 1
    wo = (w + 2 * pad - dil * (k - 1) - 1) // st + 1
    gf = 2.0 * B * ho * wo * cout * cin * k * k / 1e9
    print(f"{name} {pname} B={B}: {ms.value * 1000:.1f} us, {gf / ms.value:.1f} TF/s, tile {bn.value % 1000} splits {bn.value // 1000}")

main()
