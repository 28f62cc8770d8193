"""N4 on the device: pn_map_quantize / pn_map_sample through the C-ABI against the reference's numpy expressions
(nav/collect_maps.py:79-80, prediction/train_prediction_model.py:66-84) as restated - and checked field by field against the
reference reader's own output - in peanut_b200/map_dataset.py / tests/golden/map_dataset.npz.  Byte and integer work: bit-exact."""
import os

import numpy as np
import pytest
import torch

from peanut_b200 import map_dataset as D

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "map_dataset.npz"))


def test_sample_matches_reference_fixture(tmp_path):
    path = str(tmp_path / "f00001.npz")
    np.savez_compressed(path, maps=GOLD["maps"])
    seq = D.DeviceMapSequence.from_file(path)
    assert (seq.T, seq.C, seq.W, seq.H) == GOLD["maps"].shape
    for t_idx in (0, 3, 9, -1):
        host = D.load_map_sample("f00001.npz", t_idx, img_prefix=str(tmp_path))
        dev = seq.sample(t_idx)
        torch.cuda.synchronize()
        assert dev["img"].dtype == torch.float32 and dev["gt_semantic_seg"].dtype == torch.int64
        assert np.array_equal(dev["img"].cpu().numpy(), host["img"])
        assert np.array_equal(dev["input"].cpu().numpy(), D.network_input(host))
        assert np.array_equal(dev["gt_semantic_seg"].cpu().numpy(), host["gt_semantic_seg"])
        if t_idx >= 0:
            assert np.array_equal(dev["gt_semantic_seg"].cpu().numpy(), GOLD[f"gt_{t_idx}"])
            assert float(dev["img"].cpu().numpy().sum(dtype=np.float64)) == float(GOLD[f"img_sum_{t_idx}"])
    only_input = seq.sample(2, hwc=False, target=False)
    assert only_input["img"] is None and only_input["gt_semantic_seg"] is None and only_input["input"].shape == (14, 48, 48)


@pytest.mark.parametrize("shape", [(3, 10, 7, 7), (2, 14, 10, 26), (4, 14, 33, 64), (2, 12, 5, 3)])
def test_sample_matches_numpy_on_odd_shapes(shape):
    """Both sample kernels against the reference's numpy expressions (train_prediction_model.py:66-84): cell counts that are not
    a multiple of four take the one-cell-per-thread kernel, the others the four-cells-per-thread one (with a partial last CTA)."""
    rng = np.random.default_rng(sum(shape))
    maps = rng.integers(0, 256, shape, dtype=np.uint8)
    maps[:, 1][rng.random(maps[:, 1].shape) < 0.5] = 0
    seq = D.DeviceMapSequence(maps)
    for t_idx in (0, shape[0] - 1):
        img = maps[t_idx].transpose(1, 2, 0).astype(np.float32) / 255.
        mask = img[:, :, 1] > 0
        gt = (maps[-1, range(4, 4 + D.NUM_TARGET_CATEGORIES)] * (1 - mask)).transpose(1, 2, 0)
        dev = seq.sample(t_idx)
        torch.cuda.synchronize()
        assert np.array_equal(dev["img"].cpu().numpy(), img)
        assert np.array_equal(dev["input"].cpu().numpy(), img.transpose(2, 0, 1))
        assert np.array_equal(dev["gt_semantic_seg"].cpu().numpy(), gt) and gt.dtype == np.int64
        part = seq.sample(t_idx, hwc=False)
        assert np.array_equal(part["gt_semantic_seg"].cpu().numpy(), gt) and np.array_equal(part["input"].cpu().numpy(), img.transpose(2, 0, 1))


@pytest.mark.parametrize("shape", [(14, 32, 32), (3, 7, 11), (1, 1, 1), (14, 480, 480)])
def test_quantize_matches_numpy(shape):
    rng = np.random.default_rng(sum(shape))
    full = rng.random(shape).astype(np.float32)
    full[full > 0.98] = 1.0
    full[full < 0.02] = 0.0
    ref = (full * 255).astype(np.uint8)
    q = D.quantize_full_map(torch.from_numpy(full).cuda())
    assert q.dtype == np.uint8 and np.array_equal(q, ref)
    # a view that starts off a 16-byte boundary takes the scalar path
    flat = torch.from_numpy(full).cuda().reshape(-1)
    if flat.numel() > 1:
        assert np.array_equal(D.quantize_full_map_device(flat[1:]).cpu().numpy(), ref.reshape(-1)[1:])


def test_full_size_round_trip_and_errors():
    """At the reference's full size (20 steps x 14 channels x 960 x 960): every uint8 value survives quantise(sample(.)), the
    target is empty wherever the input has explored, and equals the last step elsewhere."""
    g = torch.Generator(device="cuda").manual_seed(5)
    maps = torch.randint(0, 256, (20, 14, 960, 960), generator=g, device="cuda", dtype=torch.uint8)
    maps[:, 1] = torch.where(maps[:, 1] > 128, maps[:, 1], torch.zeros_like(maps[:, 1]))   # about half explored
    seq = D.DeviceMapSequence(maps)
    s = seq.sample(7)
    back = D.quantize_full_map_device(s["input"] * 1.0)
    # (v / 255) * 255 truncates back to v for every uint8 v except where the fp32 product lands just below the integer
    v = torch.arange(256, dtype=torch.float32)
    lut = ((v / 255.) * 255).to(torch.uint8).cuda()
    assert torch.equal(back, lut[maps[7].long()])
    explored = maps[7, 1] > 0
    gt = s["gt_semantic_seg"]
    assert int(gt[explored].abs().sum()) == 0
    assert torch.equal(gt[~explored], maps[-1, 4:10].permute(1, 2, 0)[~explored].long())
    assert torch.equal(s["img"], s["input"].permute(1, 2, 0))
    with pytest.raises(RuntimeError):
        seq.sample(20)
    with pytest.raises(TypeError):
        D.DeviceMapSequence(torch.zeros((2, 14, 8, 8)))
