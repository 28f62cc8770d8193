"""Parity of the tcgen05 implicit-GEMM convolution (pn_conv2d) against torch.nn.functional.conv2d.

bf16 path: inputs/weights are pre-rounded to bf16 so both sides multiply identical operands; the only
differences left are fp32 accumulation order and the final bf16 rounding of the stored output
(relative 2^-8) -> tolerance 1e-2 relative to the output scale.  tf32 path: 10-bit mantissa operands.
"""
import pytest
import torch
import torch.nn.functional as F

from peanut_b200 import _lib
from tests.helpers import bf16_round, conv2d_cabi

pytestmark = pytest.mark.gpu

# (B, Cin, H, W, Cout, k, stride, dil, pad, force_bn)
CASES = [
    (1, 64, 16, 16, 64, 1, 1, 1, 0, 0),      # 1x1, tiled-A path, K = one block
    (1, 256, 20, 24, 128, 1, 1, 1, 0, 0),    # 1x1, several K blocks, M tail (480 rows)
    (2, 64, 15, 17, 64, 3, 1, 1, 1, 0),      # 3x3 pad 1, im2col, odd sizes, tile crosses images
    (1, 128, 31, 29, 128, 3, 2, 1, 1, 0),    # 3x3 stride 2 (layer2.0.conv2)
    (1, 256, 23, 23, 256, 3, 1, 2, 2, 0),    # dilation 2 (layer3)
    (1, 512, 19, 19, 512, 3, 1, 4, 4, 0),    # dilation 4 (layer4)
    (1, 256, 24, 24, 512, 1, 2, 1, 0, 0),    # 1x1 stride 2 downsample (im2col with stride)
    (1, 14, 40, 40, 32, 3, 2, 1, 1, 0),      # stem.0: 14 -> pad 16 channels, 32-byte swizzle rows
    (1, 32, 33, 35, 64, 3, 1, 1, 1, 0),      # stem.6: 32 channels, 64-byte rows
    (1, 3, 64, 48, 64, 7, 2, 1, 3, 0),       # Mask-RCNN stem 7x7 s2 p3
    (1, 512, 12, 12, 6, 1, 1, 1, 0, 0),      # classifier: Cout 6 -> N tile 32, 8 stored channels
    (1, 128, 30, 30, 256, 3, 1, 1, 1, 256),  # N tile 256
    (1, 128, 30, 30, 256, 3, 1, 1, 1, 64),   # N tile 64
    (3, 2048, 6, 6, 512, 1, 1, 1, 0, 0),     # PPM 1x1 on pooled bins, long K
    (1, 1024, 9, 9, 512, 3, 1, 1, 1, 0),     # long K loop (144 blocks) > pipeline depth
    (4, 64, 48, 48, 256, 1, 1, 1, 0, 0),     # many tiles per CTA? (72 m-tiles x n) persistent loop
    (8, 256, 40, 40, 256, 3, 1, 1, 1, 128),  # 100 m-tiles x 2 n-tiles > 148 CTAs: multi-tile persistence
    (2, 64, 15, 17, 64, 3, 1, 1, 1, 0x1000),       # direct-store epilogue (fallback path)
    (1, 256, 24, 24, 512, 1, 2, 1, 0, 0x1000 | 128),
    (1, 64, 36, 36, 48, 1, 1, 1, 0, 0),      # Cout 48: partial 128-byte chunk clipped by the TMA store
    (6, 512, 30, 30, 2048, 1, 1, 1, 0, 0),   # conv3-like: many chunks per tile, residual ring wraps
    # split-K (force_bn bits 16..23 = splits): a cluster of `splits` CTAs per tile, reduce-scatter over distributed smem
    (1, 1024, 9, 9, 512, 3, 1, 1, 1, (4 << 16) | 128),   # 144 K blocks in 4 splits, 3x3 taps cross split boundaries
    (3, 1856, 6, 6, 512, 1, 1, 1, 0, (4 << 16) | 64),    # 1x1, 29 (bf16) / 58 (tf32) K blocks in 4 uneven splits
    (2, 512, 25, 34, 512, 3, 1, 1, 1, (8 << 16) | 256),  # res5.conv2 at B=2: 14 m-tiles x 2 n-tiles x 8 splits
    (1, 256, 10, 10, 24, 3, 1, 1, 1, (2 << 16) | 32),    # narrow output (24 -> 32-wide tile), 16 columns per rank
]
SPLIT_CASES = [c for c in CASES if c[9] >> 16]
# two CTAs per SM (force_bn bits 24..25 = 1): half-size shared-memory footprint (one staging buffer, two residual buffers,
# 64-byte chunks in bf16), grid of 2 x SM count; 0x4000 | 1 << 16 = single CTA tiles, no split-K
TWO = (1 << 24) | 0x4000 | (1 << 16)
TWO_CTA_CASES = [
    (8, 256, 40, 40, 256, 3, 1, 1, 1, TWO | 128),    # 100 m-tiles x 2 n-tiles on 296 CTAs: one tile each, co-resident pairs
    (6, 512, 30, 30, 2048, 1, 1, 1, 0, TWO | 128),   # conv3-like with residual: 43 x 16 tiles, residual ring of two
    (8, 64, 100, 136, 64, 3, 1, 1, 1, TWO | 64),     # res2.conv2-like: 850 tiles of 64 columns, several tiles per CTA
    (4, 64, 48, 48, 256, 1, 1, 1, 0, TWO | 64),      # K = one block
]
# CTA pairs (force_bn bit 0x2000): tcgen05.mma.cta_group::2, a 256 x N tile on two SMs, weights split between the CTAs
PAIR = 0x2000
PAIR_CASES = [
    (2, 128, 30, 30, 256, 3, 1, 1, 1, PAIR | 256),   # 15 m-tiles (odd): the last pair's second CTA is out of range
    (1, 256, 24, 24, 512, 1, 1, 1, 0, PAIR | 256),   # 1x1 (tiled A), 5 m-tiles, 2 n-tiles
    (8, 256, 40, 40, 256, 3, 1, 1, 1, PAIR | 128),   # 50 pairs x 2 n-tiles > 74 clusters: persistent pairs
    (6, 512, 30, 30, 2048, 1, 1, 1, 0, PAIR | 256),  # conv3-like: many chunks per tile, residual ring wraps
    (1, 512, 19, 19, 512, 3, 1, 4, 4, PAIR | 256),   # dilation 4, 3 m-tiles
    (1, 256, 24, 24, 512, 1, 2, 1, 0, PAIR | 128),   # 1x1 stride 2 (im2col with stride), 2 m-tiles
    # split-K over a cluster of SM pairs (2 x splits CTAs): same reduce-scatter, peers = the same 128-row half of every pair
    (1, 256, 50, 68, 256, 3, 1, 1, 1, PAIR | (4 << 16) | 256),   # res4.conv2 at the 800 x 1088 input: 14 pairs x 4 splits
    (2, 1024, 13, 17, 512, 1, 1, 1, 0, PAIR | (2 << 16) | 128),  # 1x1, 4 m-tiles (last one partial), 4 n-tiles, 2 splits
    (1, 512, 19, 19, 512, 3, 1, 2, 2, PAIR | (4 << 16) | 128),   # dilated, 3 m-tiles: the last pair's second CTA has no rows
]


# weights resident (force_bn bits 24..26 = 4, or 5 with two CTAs per SM): persistent CTAs on one n tile each keep that tile's
# whole weight slab in shared memory, the operand ring carries activations only
RES = (4 << 24) | 0x4000 | (1 << 16)
RES2 = (5 << 24) | 0x4000 | (1 << 16)
RESIDENT_CASES = [
    (8, 64, 100, 136, 64, 3, 1, 1, 1, RES | 64),      # res2.conv2-like: nine K blocks resident, 850 tiles on 148 CTAs
    (4, 128, 48, 48, 512, 1, 1, 1, 0, RES | 128),     # 1x1, four n tiles: grid rounded to a multiple of four
    (3, 64, 97, 131, 192, 1, 1, 1, 0, RES | 64),      # three n tiles (148 -> 147 CTAs), ragged last m tile
    (16, 64, 100, 136, 256, 1, 1, 1, 0, RES2 | 128),  # res2.conv3-like with two CTAs per SM: 3 400 tiles on 296 CTAs
    (8, 64, 100, 136, 64, 1, 1, 1, 0, RES2 | 64),     # one n tile, one K block
]


def _run(ctx, case, precision, with_res=True):
    B, Cin, H, W, Cout, k, stride, dil, pad, bn = case
    g = torch.Generator().manual_seed(sum(case[:9]) + (case[9] & 0x3ff))  # same data for every launch mode of a tile width
    x = torch.randn((B, Cin, H, W), generator=g)
    w = torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5
    bias = torch.randn(Cout, generator=g) * 0.1
    if precision == _lib.PN_BF16:
        x, w = bf16_round(x), bf16_round(w)
    xd = x.cuda()
    ref = F.conv2d(xd.double(), w.cuda().double(), stride=stride, padding=pad, dilation=dil)
    ref = ref * scale.cuda().double()[None, :, None, None] + bias.cuda().double()[None, :, None, None]
    res = None
    if with_res:
        res = torch.randn(ref.shape, generator=g)
        if precision == _lib.PN_BF16:
            res = bf16_round(res)
        res = res.cuda()
        ref = ref + res.double()
    ref = torch.relu(ref).float()
    y = conv2d_cabi(ctx, xd, w, scale, bias, res, stride=stride, dil=dil, pad=pad, relu=True, force_bn=bn,
                    precision=precision)
    torch.cuda.synchronize()
    return y, ref


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_conv_bf16(ctx, case):
    y, ref = _run(ctx, case, _lib.PN_BF16)
    err = (y - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= 1e-2 * scale, f"max abs err {err} vs scale {scale}"
    # mean error far below the bf16 output rounding bound => no systematically wrong taps / channels
    assert (y - ref).abs().mean().item() <= 2e-3 * scale


@pytest.mark.parametrize("case", CASES[:12] + SPLIT_CASES, ids=[str(c) for c in CASES[:12] + SPLIT_CASES])
def test_conv_tf32(ctx, case):
    y, ref = _run(ctx, case, _lib.PN_TF32)
    err = (y - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= 3e-3 * scale, f"max abs err {err} vs scale {scale}"


FP32_CASES = CASES[:17] + SPLIT_CASES[:2]


@pytest.mark.parametrize("case", FP32_CASES, ids=[str(c) for c in FP32_CASES])
def test_conv_fp32(ctx, case):
    """Strict-parity mode: three tf32 products of hi / lo operand halves, fp32 accumulate, nothing rounded on the way out.
    Against the float64 convolution of the SAME fp32 inputs.  Measured (tools/fp32_parity_probe.py, profiles/r02_fp32_parity.txt):
    3e-7 .. 2e-6 of the output scale up to K = 2 304, 6.5e-6 at K = 9 216, 1.4e-5 with that K split four ways - against
    3e-4 .. 6e-4 on the tf32 path; the growth with K is the tensor core's fp32 accumulation, not the operand split."""
    y, ref = _run(ctx, case, _lib.PN_FP32)
    err = (y - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= 5e-5 * scale, f"max abs err {err} vs scale {scale}"


@pytest.mark.parametrize("precision", [_lib.PN_BF16, _lib.PN_TF32], ids=["bf16", "tf32"])
@pytest.mark.parametrize("case", PAIR_CASES, ids=[str(c) for c in PAIR_CASES])
def test_conv_pair(ctx, case, precision):
    y, ref = _run(ctx, case, precision)
    err = (y - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= (1e-2 if precision == _lib.PN_BF16 else 3e-3) * scale, f"max abs err {err} vs scale {scale}"
    # and bit-identical to the single-CTA kernel (same K order, same epilogue arithmetic)
    # pairs forbidden, same split-K factor (1 = none): the K order and the reduction order are then identical
    single = (case[:9] + ((case[9] & 0x3ff) | 0x4000 | (max(case[9] >> 16, 1) << 16),))
    y1, _ = _run(ctx, single, precision)
    assert torch.equal(y, y1)


@pytest.mark.parametrize("precision", [_lib.PN_BF16, _lib.PN_TF32], ids=["bf16", "tf32"])
@pytest.mark.parametrize("case", TWO_CTA_CASES, ids=[str(c) for c in TWO_CTA_CASES])
def test_conv_two_ctas_per_sm(ctx, case, precision):
    if precision == _lib.PN_TF32 and (case[9] & 0x3ff) == 128:
        pytest.skip("tf32 staging buffers + two 32 KB stages exceed half an SM's shared memory at 128 columns with a residual")
    y, ref = _run(ctx, case, precision)
    err = (y - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= (1e-2 if precision == _lib.PN_BF16 else 3e-3) * scale, f"max abs err {err} vs scale {scale}"
    # same tiles, same K order, same epilogue arithmetic as the one-CTA-per-SM launch: bit-identical
    y1, _ = _run(ctx, case[:9] + ((case[9] & 0x3ff) | 0x4000 | (1 << 16),), precision)
    assert torch.equal(y, y1)


@pytest.mark.parametrize("case", RESIDENT_CASES, ids=[str(c) for c in RESIDENT_CASES])
def test_conv_weights_resident(ctx, case):
    y, ref = _run(ctx, case, _lib.PN_BF16)
    err = (y - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= 1e-2 * scale, f"max abs err {err} vs scale {scale}"
    # same tiles, same K order, same epilogue as the launch that streams the weights: bit-identical
    plain = case[:9] + ((case[9] & 0x3ff) | 0x4000 | (1 << 16) | (case[9] & (1 << 24)),)
    y1, _ = _run(ctx, plain, _lib.PN_BF16)
    assert torch.equal(y, y1)


@pytest.mark.parametrize("precision", [_lib.PN_BF16, _lib.PN_TF32])
def test_split_k_is_bit_reproducible(ctx, precision):
    """The cluster reduction sums the partial tiles in a fixed order: repeated runs are bit-identical."""
    for case in SPLIT_CASES[:3]:
        a1, _ = _run(ctx, case, precision)
        a2, _ = _run(ctx, case, precision)
        assert torch.equal(a1, a2)


def test_conv_no_epilogue_extras(ctx):
    """scale/bias/residual all absent, no ReLU: negative values must survive."""
    g = torch.Generator().manual_seed(7)
    x = bf16_round(torch.randn((1, 64, 10, 10), generator=g))
    w = bf16_round(torch.randn((64, 64, 3, 3), generator=g) / 24.0)
    y = conv2d_cabi(ctx, x.cuda(), w, pad=1)
    ref = F.conv2d(x.cuda(), w.cuda(), padding=1)
    assert (y - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert (y < 0).any()
