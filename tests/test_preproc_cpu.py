"""CPU: (1) the depth pre-processing oracle against golden outputs of the UNMODIFIED reference function
(tests/golden/preproc_depth.npz, made by tests/golden/make_preproc_golden.py); (2) the host-side coefficient
tables the CUDA resize kernel uses (pn_pil_bilinear_coeffs, no GPU needed) replayed in numpy against Pillow itself:
the uint8 resize must be bit-exact."""
import ctypes
import os

import numpy as np
import pytest

from oracle import maskrcnn as O
from oracle import preproc as P
from peanut_b200 import _lib

GOLD = os.path.join(os.path.dirname(__file__), "golden", "preproc_depth.npz")


def test_depth_oracle_matches_reference_golden():
    z = np.load(GOLD)
    for i, s in enumerate(z["seeds"]):
        d = P.synth_depth(int(s))
        full = P.preprocess_depth(d, 0.5, 5.0)
        assert full.dtype == np.float32
        assert np.array_equal(full[2::4, 2::4], z["depth_cm_sub"][i])
        assert float(full.astype(np.float64).sum()) == float(z["full_sum"][i])
    # both branches of the per-column rule are exercised by the fixtures
    d = P.synth_depth(0)[:, :, 0]
    frac = (d == 0).mean(axis=0)
    assert (frac > 0.9).any() and ((frac > 0) & (frac <= 0.9)).any()


def test_obs_layout():
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (480, 640, 3)).astype(np.uint8)
    sem = (rng.random((480, 640, 10)) < 0.1).astype(np.float32)
    obs = P.preprocess_obs(rgb, P.synth_depth(5), sem)
    assert obs.shape == (14, 120, 160) and obs.dtype == np.float32
    assert np.array_equal(obs[:3], rgb[2::4, 2::4].transpose(2, 0, 1).astype(np.float32))
    assert np.array_equal(obs[4:], sem[2::4, 2::4].transpose(2, 0, 1))
    assert obs[3].min() >= 50.0 and obs[3].max() <= 45050.0


def _coeffs(n_in, n_out):
    lib = _lib.load()
    bounds = np.zeros((n_out, 2), np.int32)
    kk = np.zeros((n_out, 8), np.int32)
    ks = ctypes.c_int(8)
    _lib.check(lib.pn_pil_bilinear_coeffs(n_in, n_out, bounds.ctypes.data_as(ctypes.c_void_p),
                                          kk.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ks)))
    return bounds, kk, ks.value


def _apply(img, bounds, kk, ks, axis):
    """Pillow's ImagingResampleHorizontal/Vertical_8bpc in numpy along `axis`."""
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + img.shape[1:], np.uint8)
    for i in range(bounds.shape[0]):
        x0, n = bounds[i]
        acc = np.full(img.shape[1:], 1 << 21, np.int64)
        for j in range(n):
            acc += img[x0 + j] * int(kk[i, j])
        out[i] = np.clip(acc >> 22, 0, 255)
    return np.moveaxis(out, 0, axis)


@pytest.mark.parametrize("hw,cfg", [((480, 640), O.Cfg()), ((240, 320), O.Cfg(min_size=400, max_size=667)),
                                     ((97, 131), O.Cfg(min_size=160, max_size=200))])
def test_resize_tables_replay_pillow_bit_exactly(hw, cfg):
    from PIL import Image
    h, w = hw
    img = O.synth_rgb(1, h, w)
    nh, nw = O.resized_shape(h, w, cfg)
    ref = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BILINEAR))
    hb, hk, hks = _coeffs(w, nw)
    vb, vk, vks = _coeffs(h, nh)
    tmp = _apply(img, hb, hk, hks, axis=1)      # horizontal pass first, uint8 round trip in between
    got = _apply(tmp, vb, vk, vks, axis=0)
    assert np.array_equal(got, ref)
