"""oracle/prednet.py against the golden vectors produced by the reference's own model code
(tests/golden/make_prednet_golden.py: the unmodified mmseg ResNetV1c / PSPHead / EncoderDecoder sources over a restated
mmcv shim; bit-equality with the oracle is required when the fixtures are written)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import prednet as O

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "prednet_*.npz")))


def test_fixtures_present():
    assert len(GOLD) >= 3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    C, H, W, wseed, xseed, nparams = (int(v) for v in g["meta"])
    sd = O.synth_state_dict(C, 6, seed=wseed)
    assert sum(v.numel() for k, v in sd.items() if not k.endswith("num_batches_tracked")) == nparams
    model = O.build(sd, in_channels=C)
    x = O.synth_partial_map(C, H, W, seed=xseed)
    got = O.run_inference(model, x)[0]
    assert got.shape == (6, H, W) and got.dtype == np.float32
    # bit-exact on the machine that wrote the fixture; other hosts may pick different oneDNN kernels (fp32 reassociation)
    err = float(np.abs(got - g["logits"]).max())
    assert err <= 2e-5 * float(np.abs(g["logits"]).max()), err
    with torch.no_grad():
        feats = model.backbone(torch.from_numpy(x)[None])
    assert np.allclose([float(f.abs().max()) for f in feats], g["stage_absmax"], rtol=1e-4)
    # the folded-BN form the CUDA path implements is the same function
    folded = O.forward_folded(model, torch.from_numpy(x)[None])[0].numpy()
    assert float(np.abs(folded - g["logits"]).max()) <= 1e-4 * float(np.abs(g["logits"]).max())


def test_structure_census():
    """61 convolutions; 46.61 M parameters at 14 input channels (backbone + decode head)."""
    m = O.EncoderDecoder(14, 6)
    assert sum(1 for x in m.modules() if isinstance(x, torch.nn.Conv2d)) == 61
    n = sum(p.numel() for p in m.parameters())
    assert abs(n - 46.61e6) < 0.02e6, n


def test_get_prediction_is_expit():
    from scipy.special import expit
    sd = O.synth_state_dict(14, 6, seed=0)
    model = O.build(sd)
    x = O.synth_partial_map(14, 32, 32, seed=1)
    assert np.array_equal(O.get_prediction(model, x), expit(O.run_inference(model, x)[0]))
