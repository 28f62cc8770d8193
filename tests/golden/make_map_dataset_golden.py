"""Checks peanut_b200/map_dataset.py against the reference's own reader and writer code and writes a small fixture.

    python tests/golden/make_map_dataset_golden.py        (build container only: reads /root/reference)

Reader: the UNMODIFIED class ``LoadMapFromFile`` and the method ``SemMapDataset.load_annotations`` are cut out of
prediction/train_prediction_model.py with ``ast`` (the module imports mmcv / mmseg registries, absent here) and run on a
synthetic ``.npz`` written in the reference's format; every field they produce must equal ``load_map_sample`` /
``list_samples`` in value and dtype.  Writer: collect_maps.py is a script around Habitat, so its three statements
(:79-80 quantisation, :84 save condition, :85 savez) are restated in the checker below.
"""
import ast
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from peanut_b200 import map_dataset as D  # noqa: E402

SRC = "/root/reference/prediction/train_prediction_model.py"


def synth_sequence(seed, T=20, C=14, n=48):
    rng = np.random.default_rng(seed)
    seq = np.zeros((T, C, n, n), np.float32)
    explored = np.zeros((n, n), bool)
    for t in range(T):
        explored |= rng.random((n, n)) < 0.06
        seq[t, 1] = explored * rng.random((n, n))
        seq[t, 0] = (rng.random((n, n)) < 0.05) * explored
        for c in range(4, C):
            seq[t, c] = np.maximum(seq[t - 1, c] if t else 0, (rng.random((n, n)) < 0.01) * explored * rng.random((n, n)))
    return seq.astype(np.float32)


def reference_objects():
    tree = ast.parse(open(SRC).read())
    loader = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "LoadMapFromFile")
    loader.decorator_list = []  # @PIPELINES.register_module(): registry bookkeeping only
    dataset = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SemMapDataset")
    load_ann = next(n for n in dataset.body if isinstance(n, ast.FunctionDef) and n.name == "load_annotations")
    consts = [n for n in tree.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "").startswith("NUM_")]
    mmcv = types.SimpleNamespace(FileClient=lambda **kw: object())
    ns = {"np": np, "osp": os.path, "mmcv": mmcv, "print_log": lambda *a, **k: None, "get_root_logger": lambda: None}
    exec(compile(ast.Module(body=consts + [loader, load_ann], type_ignores=[]), SRC, "exec"), ns)
    return ns["LoadMapFromFile"], ns["load_annotations"]


def main():
    Loader, load_annotations = reference_objects()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "sub"))
        for k, (seed, rel) in enumerate([(1, "f00001.npz"), (2, "sub/f00007.npz")]):
            seq = synth_sequence(seed)
            w = D.MapSequenceWriter(seq.shape[1], seq.shape[2], seq.shape[3])
            step = 0
            for t in range(seq.shape[0]):
                for _ in range(25):  # 25 environment steps between samples
                    step += 1
                    w.record(step, seq[t])
            # restated writer statements (collect_maps.py:79-80, :84)
            want_seq = np.stack([(m * 255).astype(np.uint8) for m in seq])
            assert np.array_equal(w.seq, want_seq) and w.seq.dtype == np.uint8
            assert w.should_save() == bool(np.sum(want_seq[:, 4:]) > 0 and np.sum(want_seq[:, 1]) > 4000)
            assert w.save(os.path.join(tmp, rel))
            if k == 0:
                out["maps"] = w.seq
        # reader
        for rel in ("f00001.npz", "sub/f00007.npz"):
            for t_idx in (0, 3, 9):
                results = {"img_prefix": tmp, "img_info": {"filename": rel, "t_idx": t_idx}, "seg_fields": []}
                ref = Loader()(results)
                got = D.load_map_sample(rel, t_idx, img_prefix=tmp)
                for key, val in got.items():
                    r = ref[key]
                    if isinstance(val, np.ndarray):
                        assert r.dtype == val.dtype and np.array_equal(r, val), (rel, t_idx, key, r.dtype, val.dtype)
                    elif isinstance(val, dict):
                        assert np.array_equal(r["mean"], val["mean"]) and np.array_equal(r["std"], val["std"]) and r["to_rgb"] == val["to_rgb"]
                    else:
                        assert r == val, (key, r, val)
                assert ref["seg_fields"] == ["gt_semantic_seg"]
                if rel == "f00001.npz":
                    out[f"img_sum_{t_idx}"] = np.array(got["img"].sum(dtype=np.float64))
                    out[f"gt_{t_idx}"] = got["gt_semantic_seg"].astype(np.uint8)
        stub = types.SimpleNamespace(file_client=types.SimpleNamespace(
            list_dir_or_file=lambda dir_path, list_dir, suffix, recursive: ["sub/f00007.npz", "f00001.npz"]))
        ref_infos = load_annotations(stub, tmp, ".npz", None, ".npz", None)
        assert ref_infos == D.list_samples(tmp), "sample enumeration differs"
        print(f"reader fields, dtypes and the sample enumeration ({len(ref_infos)} entries) equal the reference's")
    np.savez_compressed(os.path.join(HERE, "map_dataset.npz"), **out)


if __name__ == "__main__":
    main()
