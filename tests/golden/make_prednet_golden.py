"""Pins oracle/prednet.py to the reference's own model code and writes tests/golden/prednet_*.npz.

Run in the build container (needs /root/reference; the GPU box and the test-suite only read the committed .npz files):

    python tests/golden/make_prednet_golden.py

What executes here is the UNMODIFIED reference source — `prediction/mmseg/models/backbones/resnet.py` (ResNetV1c, Bottleneck),
`models/utils/res_layer.py`, `models/decode_heads/{decode_head,psp_head,fcn_head}.py`, `models/segmentors/{base,encoder_decoder}.py`,
`ops/wrappers.py`, `models/builder.py` — built from the reference's own config `nav/pred_model_cfg.py` through
`builder.build_segmentor`, run through `EncoderDecoder.simple_test(..., rescale=True)` exactly as `nav/agent/prediction.py:112-137`
does (its test pipeline is a numerical identity: `MapFromArray` + `ImageToTensor`, no resize / flip / normalisation at ratio 1).
The only thing that is NOT the reference is mmcv-full 1.6.0 (absent, un-vendored): its conv/norm builders, ConvModule, BaseModule
and Registry are restated in mmcv_shim.py.  The package `__init__` files of mmseg are bypassed (they pull in every backbone/head of
the model zoo and assert the mmcv version); modules are imported file by file from the reference tree.

The script REQUIRES bit-equality between the reference model and oracle/prednet.py on every case before it writes a fixture, and
checks that the oracle's synthetic checkpoint loads into the reference model with no unexpected / missing backbone or decode-head key
(so the key naming of a real `pred_model_wts.pth` is honoured).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import mmcv_shim  # noqa: E402
from oracle import prednet as O  # noqa: E402


def _skeleton(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    if "." in name:
        setattr(sys.modules[name.rsplit(".", 1)[0]], name.rsplit(".", 1)[1], m)
    return m


def load_reference_modules():
    """Import the reference's model files without running mmseg's package __init__ files."""
    mmcv_shim.install()
    base = os.path.join(REF, "prediction", "mmseg")
    _skeleton("mmseg", base)
    _skeleton("mmseg.ops", os.path.join(base, "ops"))
    core = _skeleton("mmseg.core", os.path.join(base, "core"))
    core.build_pixel_sampler = lambda cfg, **kw: None          # sampler=None on this config (decode_head.py:96-99)
    core.add_prefix = lambda d, p: {f"{p}.{k}": v for k, v in d.items()}
    _skeleton("mmseg.models", os.path.join(base, "models"))
    for sub in ("backbones", "decode_heads", "segmentors", "utils", "losses"):
        _skeleton("mmseg.models." + sub, os.path.join(base, "models", sub))
    wrappers = importlib.import_module("mmseg.ops.wrappers")
    sys.modules["mmseg.ops"].resize = wrappers.resize
    sys.modules["mmseg.ops"].Upsample = wrappers.Upsample
    sys.modules["mmseg.models.losses"].accuracy = lambda *a, **k: None   # training-only metric
    builder = importlib.import_module("mmseg.models.builder")
    res_layer = importlib.import_module("mmseg.models.utils.res_layer")
    sys.modules["mmseg.models.utils"].ResLayer = res_layer.ResLayer
    importlib.import_module("mmseg.models.backbones.resnet")
    importlib.import_module("mmseg.models.decode_heads.psp_head")
    importlib.import_module("mmseg.models.decode_heads.fcn_head")
    importlib.import_module("mmseg.models.segmentors.encoder_decoder")

    @builder.LOSSES.register_module(name="MyLoss")          # nav/agent/prediction.py:71-108; training only
    class MyLoss(torch.nn.Module):
        def __init__(self, loss_weight=1.0, **kw):
            super().__init__()
            self.loss_weight = loss_weight

    return builder


class _Cfg(dict):
    """mmcv.Config-style attribute access for the test_cfg / train_cfg dicts the segmentor reads."""
    __getattr__ = dict.get


def reference_model_cfg(in_channels):
    ns = {}
    exec(open(os.path.join(REF, "nav", "pred_model_cfg.py")).read(), ns)
    cfg = ns["model"]
    assert cfg["backbone"]["in_channels"] == 14 and cfg["decode_head"]["num_classes"] == 6
    cfg["backbone"]["in_channels"] = in_channels      # BASELINE.json's 24-channel variant reuses the same architecture
    cfg["pretrained"] = None                          # init_segmentor does the same (mmseg/apis/inference.py:29)
    cfg["backbone"].pop("pretrained", None)
    cfg["train_cfg"] = None
    cfg["test_cfg"] = _Cfg(cfg["test_cfg"])
    return cfg


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    builder = load_reference_modules()
    # the last case is BASELINE.json's headline map shape (24 x 240 x 240) with the weights bench.py uses (seed 0)
    cases = [("c14_96", 14, 96, 96, 0, 3), ("c14_120x88", 14, 120, 88, 1, 5), ("c24_64", 24, 64, 64, 2, 7),
             ("c24_240", 24, 240, 240, 0, 1234)]
    for name, C, H, W, wseed, xseed in cases:
        ref = builder.build_segmentor(reference_model_cfg(C))
        sd = O.synth_state_dict(C, 6, seed=wseed)
        res = ref.load_state_dict(sd, strict=False)
        assert not res.unexpected_keys, res.unexpected_keys
        assert all(k.startswith("auxiliary_head.") for k in res.missing_keys), res.missing_keys
        ref.eval()
        x = O.synth_partial_map(C, H, W, seed=xseed)
        img = torch.from_numpy(x)[None]
        meta = [dict(ori_shape=(H, W, C), img_shape=(H, W, C), pad_shape=(H, W, C), scale_factor=1.0, flip=False)]
        with torch.no_grad():
            got_ref = ref(img=[img], img_metas=[meta], return_loss=False, rescale=True)[0]   # prediction.py:135-136
            feats = ref.extract_feat(img)
        oracle = O.build(sd, in_channels=C)
        got_orc = O.run_inference(oracle, x)[0]
        with torch.no_grad():
            feats_o = oracle.backbone(img)
        for a, b in zip(feats, feats_o):
            assert torch.equal(a, b), f"{name}: backbone stage output differs from the reference"
        assert got_ref.dtype == np.float32 and got_ref.shape == (6, H, W)
        assert np.array_equal(got_ref, got_orc), f"{name}: oracle != reference (max {np.abs(got_ref - got_orc).max()})"
        prob = O.get_prediction(oracle, x)
        from scipy.special import expit                                                   # prediction.py:22-23,158
        assert prob.dtype == np.float32 and np.array_equal(prob, expit(got_ref))
        nparams = sum(p.numel() for k, p in ref.state_dict().items() if not k.startswith("auxiliary_head")
                      and not k.endswith("num_batches_tracked"))
        np.savez_compressed(os.path.join(HERE, f"prednet_{name}.npz"), logits=got_ref.astype(np.float32),
                            stage_absmax=np.array([float(f.abs().max()) for f in feats], np.float32),
                            stage_mean=np.array([float(f.mean()) for f in feats], np.float32),
                            meta=np.array([C, H, W, wseed, xseed, nparams], np.int64))
        print(f"{name}: reference == oracle bit-exact on {got_ref.shape}, logits range [{got_ref.min():.3f}, {got_ref.max():.3f}], "
              f"{nparams} tensors' elements")


if __name__ == "__main__":
    main()
