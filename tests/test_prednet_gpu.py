"""Stage C parity: CUDA map-completion network (through the reference-facing shim and the C-ABI)
against the CPU oracle (oracle/prednet.py) on the same seeded weights and inputs.

Tolerances (stated, evidence in DESIGN.md §5):
  * bf16 path vs the bf16-storage emulation of the oracle: 2e-2 of the logit range (accumulation order +
    single-ulp bf16 flips that propagate through 50 layers);
  * bf16 path vs the fp32 oracle: 5e-2 of the logit range on logits, 1.5e-2 abs on probabilities;
  * tf32 path vs the fp32 oracle: 5e-3 of the logit range, 1.5e-3 abs on probabilities.
"""
import numpy as np
import pytest
import torch

from oracle import prednet as oracle
from peanut_b200 import prediction

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def weights():
    return oracle.synth_state_dict(in_channels=14, num_classes=6, seed=0)


@pytest.fixture(scope="module")
def oracle_model(weights):
    return oracle.build(weights)


def _segmentor(weights, precision):
    return prediction.init_segmentor(prediction._default_cfg(14, 6), device="cuda:0", precision=precision,
                                     state_dict=weights)


@pytest.mark.parametrize("hw", [(64, 64), (240, 240), (200, 136)])
def test_bf16_vs_emulated_oracle(weights, oracle_model, hw):
    H, W = hw
    x = oracle.synth_partial_map(14, H, W, seed=11)
    seg = _segmentor(weights, "bf16")
    got = seg.forward_device(torch.from_numpy(x)[None].cuda())
    torch.cuda.synchronize()
    ref, feats, lowres = oracle.forward_folded(oracle_model, torch.from_numpy(x)[None], emulate_bf16=True,
                                               return_taps=True)
    f_got = seg.read_tap(0).cpu()
    rng_f = feats.abs().max().item()
    assert (f_got - feats).abs().max().item() <= 4e-2 * rng_f, "layer4 features"
    l_got = seg.read_tap(1).cpu()
    rng = lowres.abs().max().item()
    assert (l_got - lowres).abs().max().item() <= 2e-2 * rng, "low-res logits"
    assert (got.cpu() - ref).abs().max().item() <= 2e-2 * rng, "full-res logits"


def test_bf16_vs_fp32_oracle_reference_api(weights, oracle_model):
    """Through PEANUT_Prediction_Model.get_prediction / run_inference with host numpy in and out."""
    import types
    x = oracle.synth_partial_map(14, 240, 240, seed=5)
    args = types.SimpleNamespace(sem_gpu_id=0, pred_model_wts=None, pred_model_cfg=None)
    model = prediction.PEANUT_Prediction_Model(args, state_dict=weights, precision="bf16")
    logits = prediction.run_inference(model.model, x)
    assert isinstance(logits, list) and logits[0].shape == (6, 240, 240) and logits[0].dtype == np.float32
    ref_logits = oracle.run_inference(oracle_model, x)[0]
    rng = np.abs(ref_logits).max()
    assert np.abs(logits[0] - ref_logits).max() <= 5e-2 * rng
    prob = model.get_prediction(x)
    assert prob.flags.writeable and prob.shape == (6, 240, 240)
    ref_prob = oracle.get_prediction(oracle_model, x)
    assert np.abs(prob - ref_prob).max() <= 1.5e-2
    # integer category map derived from the prediction (argmax over classes) must agree wherever the
    # oracle's top-2 margin exceeds the stated tolerance
    srt = np.sort(ref_logits, axis=0)
    confident = (srt[-1] - srt[-2]) > 2 * 5e-2 * rng
    assert (logits[0].argmax(0)[confident] == ref_logits.argmax(0)[confident]).all()


def test_tf32_vs_fp32_oracle(weights, oracle_model):
    x = oracle.synth_partial_map(14, 240, 240, seed=6)
    seg = _segmentor(weights, "tf32")
    got = seg.forward_device(torch.from_numpy(x)[None].cuda()).cpu().numpy()[0]
    ref = oracle.run_inference(oracle_model, x)[0]
    rng = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 5e-3 * rng
    p = seg.forward_device(torch.from_numpy(x)[None].cuda(), apply_sigmoid=True).cpu().numpy()[0]
    assert np.abs(p - oracle.get_prediction(oracle_model, x)).max() <= 1.5e-3


def test_fp32_vs_fp32_oracle(weights, oracle_model):
    """Strict-parity mode (three tf32 products per convolution, no rounding of stored activations) against the fp32 oracle.
    Measured (tools/fp32_parity_probe.py, profiles/r02_fp32_parity.txt): logits 1.06e-4 of their range (tf32 path: 1.49e-3),
    probabilities 1.9e-5 (2.7e-4), 99.9965 % of the cells with the oracle's class (99.90 %).  The CPU oracle itself sits 1e-6
    from a float64 evaluation; what is left here is the tensor core's fp32 accumulation over 50 layers."""
    x = oracle.synth_partial_map(14, 240, 240, seed=6)
    seg = _segmentor(weights, "fp32")
    got = seg.forward_device(torch.from_numpy(x)[None].cuda()).cpu().numpy()[0]
    ref = oracle.run_inference(oracle_model, x)[0]
    rng = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 3e-4 * rng
    p = seg.forward_device(torch.from_numpy(x)[None].cuda(), apply_sigmoid=True).cpu().numpy()[0]
    assert np.abs(p - oracle.get_prediction(oracle_model, x)).max() <= 6e-5
    assert (got.argmax(0) == ref.argmax(0)).mean() >= 0.9998
    srt = np.sort(ref, axis=0)
    confident = (srt[-1] - srt[-2]) > 2 * 3e-4 * rng
    assert (got.argmax(0)[confident] == ref.argmax(0)[confident]).all()


def test_batch_and_graph_replay(weights, oracle_model):
    """B=3: repeated calls (CUDA-graph replay) are bit-stable; per-sample results agree with a B=1 engine up to the launch
    configuration (the measured tile / split-K choice depends on the batch's tile count, so fp32 sums are associated
    differently and bf16 roundings may flip)."""
    xs = np.stack([oracle.synth_partial_map(14, 96, 96, seed=s) for s in (1, 2, 3)])
    seg = _segmentor(weights, "bf16")
    xd = torch.from_numpy(xs).cuda()
    a = seg.forward_device(xd).clone()
    b = seg.forward_device(xd).clone()
    c = seg.forward_device(xd).clone()
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(b, c)
    seg1 = _segmentor(weights, "bf16")
    for i in range(3):
        one = seg1.forward_device(xd[i:i + 1])
        assert (one[0] - a[i]).abs().max().item() <= 2e-2 * a[i].abs().max().item()


def test_wrong_channels_raises(weights):
    seg = _segmentor(weights, "bf16")
    with pytest.raises(RuntimeError):
        seg.forward_device(torch.zeros((1, 13, 64, 64), device="cuda"))


def _golden_cases():
    import glob
    import os
    return sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "prednet_*.npz")))


@pytest.mark.parametrize("path", _golden_cases(), ids=lambda p: p.rsplit("/", 1)[-1])
def test_tf32_vs_reference_golden(path):
    """CUDA path against logits produced by the reference's own model code (tests/golden/make_prednet_golden.py):
    tf32 operands / fp32 accumulate -> 5e-3 of the logit range; argmax equal wherever the top-2 margin exceeds that."""
    g = np.load(path)
    C, H, W, wseed, xseed, _ = (int(v) for v in g["meta"])
    sd = oracle.synth_state_dict(C, 6, seed=wseed)
    seg = prediction.init_segmentor(prediction._default_cfg(C, 6), device="cuda:0", precision="tf32", state_dict=sd)
    x = oracle.synth_partial_map(C, H, W, seed=xseed)
    got = prediction.run_inference(seg, x)[0]
    ref = g["logits"]
    rng = float(np.abs(ref).max())
    tol = 5e-3 * rng
    assert got.shape == ref.shape and float(np.abs(got - ref).max()) <= tol
    top2 = np.sort(ref, axis=0)[-2:]
    decided = (top2[1] - top2[0]) > 2 * tol
    assert decided.mean() > 0.5
    assert np.array_equal(got.argmax(0)[decided], ref.argmax(0)[decided])


@pytest.mark.parametrize("path", [p for p in _golden_cases() if "c24_" in p], ids=lambda p: p.rsplit("/", 1)[-1])
def test_bf16_vs_reference_golden_24_channels(path):
    """The throughput path (bf16 storage) at BASELINE.json's map shape (24 x 240 x 240, and 24 x 64 x 64) against the logits
    of the reference's own model code: 5e-2 of the logit range; argmax category map equal wherever the reference's top-2
    margin exceeds twice that; probabilities within 1.5e-2."""
    g = np.load(path)
    C, H, W, wseed, xseed, _ = (int(v) for v in g["meta"])
    sd = oracle.synth_state_dict(C, 6, seed=wseed)
    seg = prediction.init_segmentor(prediction._default_cfg(C, 6), device="cuda:0", precision="bf16", state_dict=sd)
    x = oracle.synth_partial_map(C, H, W, seed=xseed)
    got = prediction.run_inference(seg, x)[0]
    ref = g["logits"]
    rng = float(np.abs(ref).max())
    tol = 5e-2 * rng
    err = float(np.abs(got - ref).max())
    assert got.shape == ref.shape and err <= tol, (err, tol)
    top2 = np.sort(ref, axis=0)[-2:]
    decided = (top2[1] - top2[0]) > 2 * tol
    assert decided.mean() > 0.2
    assert np.array_equal(got.argmax(0)[decided], ref.argmax(0)[decided])
    # the fraction of ALL cells whose argmax agrees is reported by the assertion message if it ever drops
    agree = float((got.argmax(0) == ref.argmax(0)).mean())
    assert agree >= 0.97, agree
    xd = torch.from_numpy(x)[None].cuda()
    prob = seg.forward_device(xd, apply_sigmoid=True).cpu().numpy()[0]
    assert float(np.abs(prob - 1.0 / (1.0 + np.exp(-ref.astype(np.float64)))).max()) <= 1.5e-2


def test_tf32_reference_shape_720(weights, oracle_model):
    """The reference's own geometry (14 x 720 x 720, nav/arguments.py:40,74): tf32 path vs the fp32 oracle."""
    x = oracle.synth_partial_map(14, 720, 720, seed=21)
    seg = _segmentor(weights, "tf32")
    got = prediction.run_inference(seg, x)[0]
    ref = oracle.run_inference(oracle_model, x)[0]
    rng = float(np.abs(ref).max())
    assert got.shape == (6, 720, 720)
    assert float(np.abs(got - ref).max()) <= 5e-3 * rng
