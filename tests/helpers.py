"""Shared helpers for the parity tests."""
import ctypes

import numpy as np
import torch

from peanut_b200 import _lib


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def conv2d_cabi(ctx, x, w, scale=None, bias=None, residual=None, stride=1, dil=1, pad=0, relu=False, force_bn=0,
                precision=_lib.PN_BF16):
    """Calls pn_conv2d on CUDA tensors x [B,Cin,H,W], residual [B,Cout,Ho,Wo]; w/scale/bias host tensors."""
    B, Cin, H, W = x.shape
    Cout, _, R, S = w.shape
    Ho = (H + 2 * pad - dil * (R - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (S - 1) - 1) // stride + 1
    y = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device)
    wh = np.ascontiguousarray(w.cpu().numpy(), dtype=np.float32)
    sh = None if scale is None else np.ascontiguousarray(scale.cpu().numpy(), dtype=np.float32)
    bh = None if bias is None else np.ascontiguousarray(bias.cpu().numpy(), dtype=np.float32)
    p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    x = x.contiguous()
    if residual is not None:
        residual = residual.contiguous()
    _lib.check(ctx.lib.pn_conv2d(ctx.handle, precision, x.data_ptr(), B, Cin, H, W, p(wh), p(sh), p(bh),
                                 None if residual is None else residual.data_ptr(), Cout, R, S, stride, dil, pad,
                                 int(relu), force_bn, y.data_ptr()))
    return y
