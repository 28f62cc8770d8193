"""GPU diagnostic: runs each conv parity case in its own process (a trapped kernel poisons the CUDA
context) and prints error statistics with enough structure to localise a descriptor / layout bug.
Usage: python tools/conv_probe.py            (driver)
       python tools/conv_probe.py <case> <precision>   (one case)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(idx, precision):
    import torch
    import torch.nn.functional as F
    from peanut_b200 import _lib
    from tests.helpers import bf16_round, conv2d_cabi
    from tests.test_conv_gpu import CASES
    case = CASES[idx]
    B, Cin, H, W, Cout, k, stride, dil, pad, bn = case
    ctx = _lib.Context(0)
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn((B, Cin, H, W), generator=g)
    w = torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5
    if precision == 0:
        x, w = bf16_round(x), bf16_round(w)
    ref = F.conv2d(x.cuda().double(), w.cuda().double(), stride=stride, padding=pad, dilation=dil).float()
    y = conv2d_cabi(ctx, x.cuda(), w, stride=stride, dil=dil, pad=pad, force_bn=bn, precision=precision)
    torch.cuda.synchronize()
    err = (y - ref).abs()
    scale = ref.abs().max().item()
    print(f"case {idx} {case} prec={precision}: max_err={err.max().item():.4g} mean_err={err.mean().item():.4g} "
          f"scale={scale:.4g} nan={torch.isnan(y).any().item()} zero_frac={(y == 0).float().mean().item():.3f}")
    if err.max().item() > 1e-2 * scale:
        # per-channel-block and per-row-block structure
        e = err.permute(0, 2, 3, 1).reshape(-1, Cout)  # [M, Cout]
        M = e.shape[0]
        rows = [e[i:i + 32].max().item() for i in range(0, min(M, 512), 32)]
        cols = [e[:, j:j + 8].max().item() for j in range(0, min(Cout, 128), 8)]
        print("  row-block(32) max err:", " ".join(f"{v:.2g}" for v in rows))
        print("  col-block(8)  max err:", " ".join(f"{v:.2g}" for v in cols))
        yy = y.permute(0, 2, 3, 1).reshape(-1, Cout)
        rr = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
        print("  y[0,:8]  ", yy[0, :8].tolist())
        print("  ref[0,:8]", rr[0, :8].tolist())
        print("  y[1,:8]  ", yy[1, :8].tolist())
        print("  ref[1,:8]", rr[1, :8].tolist())


if __name__ == "__main__":
    if len(sys.argv) >= 3:
        one(int(sys.argv[1]), int(sys.argv[2]))
    else:
        from tests.test_conv_gpu import CASES
        for prec in (0, 1):
            for i in range(len(CASES)):
                try:
                    r = subprocess.run([sys.executable, __file__, str(i), str(prec)], capture_output=True, text=True,
                                       timeout=180)
                    out = (r.stdout + r.stderr[-1500:]) if r.returncode != 0 else r.stdout
                    print(out.strip(), flush=True)
                except subprocess.TimeoutExpired:
                    print(f"case {i} prec={prec}: TIMEOUT", flush=True)
