"""CPU: the map-sequence file format.  peanut_b200/map_dataset.py was checked field by field against the
reference's own LoadMapFromFile / SemMapDataset.load_annotations by tests/golden/make_map_dataset_golden.py, which wrote
tests/golden/map_dataset.npz (one quantised sequence, the targets and input checksums the reference reader produced)."""
import os

import numpy as np
import torch

from peanut_b200 import map_dataset as D

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "map_dataset.npz"))


def test_reader_matches_reference_fixture(tmp_path):
    path = str(tmp_path / "f00001.npz")
    np.savez_compressed(path, maps=GOLD["maps"])
    for t_idx in (0, 3, 9):
        s = D.load_map_sample("f00001.npz", t_idx, img_prefix=str(tmp_path))
        assert s["img"].dtype == np.float32 and s["img"].shape == (48, 48, 14)
        assert float(s["img"].sum(dtype=np.float64)) == float(GOLD[f"img_sum_{t_idx}"])
        assert s["gt_semantic_seg"].dtype == np.int64 and s["gt_semantic_seg"].shape == (48, 48, 6)
        assert np.array_equal(s["gt_semantic_seg"], GOLD[f"gt_{t_idx}"])
        # the target only lives where the input has not explored yet
        assert not (s["gt_semantic_seg"].sum(-1) * (s["img"][:, :, 1] > 0)).any()
        x = D.network_input(s)
        assert x.shape == (14, 48, 48) and x.flags["C_CONTIGUOUS"] and x.max() <= 1.0


def test_writer_quantisation_and_save_rule(tmp_path):
    rng = np.random.default_rng(0)
    full = rng.random((14, 32, 32)).astype(np.float32)
    full[full > 0.98] = 1.0
    q = D.quantize_full_map(full)
    assert q.dtype == np.uint8 and np.array_equal(q, (full * 255).astype(np.uint8))
    assert np.array_equal(D.quantize_full_map(torch.from_numpy(full)), q)  # tensor path (device-resident maps) == numpy path
    assert q[full == 1.0].min() == 255 and q.max() == 255

    w = D.MapSequenceWriter(14, 32, 32)
    assert len(w.save_steps) == 20 and w.save_steps[0] == 25 and w.save_steps[-1] == 500
    stored = [w.record(step, full) for step in range(1, 60)]
    assert stored.count(True) == 2 and w.seq_i == 2 and np.array_equal(w.seq[1], q)
    assert w.should_save() and w.save(str(tmp_path / "a.npz"))
    assert np.array_equal(np.load(str(tmp_path / "a.npz"))["maps"], w.seq)

    empty = D.MapSequenceWriter(14, 32, 32)
    sparse = np.zeros((14, 32, 32), np.float32)
    sparse[1, :4, :4] = 1.0   # explored sum 16 * 255 = 4080 > 4000, but nothing semantic
    empty.record(25, sparse)
    assert not empty.should_save() and not empty.save(str(tmp_path / "b.npz")) and not os.path.exists(str(tmp_path / "b.npz"))
    sparse[5, 0, 0] = 0.5
    empty.record(50, sparse)
    assert empty.should_save()


def test_sample_enumeration(tmp_path):
    os.makedirs(tmp_path / "val" / "x")
    for rel in ("val/f2.npz", "val/x/f1.npz", "val/readme.txt"):
        open(tmp_path / rel, "wb").close()
    infos = D.list_samples(str(tmp_path / "val"))
    assert len(infos) == 20 and infos[0] == {"filename": "f2.npz", "t_idx": 0} and infos[10]["filename"] == os.path.join("x", "f1.npz")
    assert [i["t_idx"] for i in infos[:10]] == list(range(10))
