"""Minimal restatement of the parts of mmcv-full 1.6.0 (peanut.Dockerfile:18; not vendored, not installable offline) that the
reference's map-completion network touches, so that the UNMODIFIED reference model files under
/root/reference/prediction/mmseg/models can be imported and executed in the build container.

Test infrastructure only (used by make_prednet_golden.py); nothing here ships or runs on the GPU box.

What is restated, and the reference call sites that depend on it:
  * mmcv.cnn.build_conv_layer(None, ...) -> nn.Conv2d              resnet.py:164-209, 595-623; res_layer.py:56-63
  * mmcv.cnn.build_norm_layer(dict(type='BN'), C, postfix) -> ('bn<postfix>', nn.BatchNorm2d(C, eps=1e-5))
                                                                    resnet.py:164-167, 598-622; res_layer.py:64
  * mmcv.cnn.ConvModule: conv (bias only when there is no norm) -> norm -> ReLU(inplace), attributes conv / bn / activate
                                                                    psp_head.py:39-46, 86-93; fcn_head.py
  * mmcv.runner.BaseModule / Sequential / ModuleList, auto_fp16 / force_fp32 (identity: fp16_enabled is False)
  * mmcv.utils.Registry (register_module / build from a cfg dict with a 'type' key)      models/builder.py
Everything else mmcv exports is a placeholder that raises if it is ever called.
"""
import sys
import types

import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm


class _Placeholder:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise RuntimeError(f"mmcv shim: {self._name} is not restated (off the inference path)")

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Placeholder(self._name + "." + item)


class _Lenient(types.ModuleType):
    """A module whose unknown attributes are placeholders (decorators / classes that are imported but never used)."""

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Placeholder(self.__name__ + "." + item)


class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self.name, self.parent, self._modules = name, parent, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self._modules[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        if key in self._modules:
            return self._modules[key]
        return self.parent.get(key) if self.parent is not None else None

    def build(self, cfg, default_args=None):
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        cls = self.get(args.pop("type"))
        assert cls is not None, f"{cfg['type']} is not registered in {self.name}"
        return cls(**args)


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = init_cfg

    @property
    def is_init(self):
        return self._is_init

    def init_weights(self):  # weights come from a state dict in every use of this shim
        self._is_init = True


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


def _identity_decorator(*dargs, **dkwargs):
    def deco(fn):
        return fn
    return deco


def build_conv_layer(cfg, *args, **kwargs):
    assert cfg is None or cfg.get("type") in ("Conv2d", "Conv"), cfg
    return nn.Conv2d(*args, **kwargs)


def build_norm_layer(cfg, num_features, postfix=""):
    cfg = dict(cfg)
    assert cfg.pop("type") == "BN"
    requires_grad = cfg.pop("requires_grad", True)
    cfg.setdefault("eps", 1e-5)
    layer = nn.BatchNorm2d(num_features, **cfg)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return "bn" + str(postfix), layer


def build_plugin_layer(*a, **k):
    raise RuntimeError("mmcv shim: plugins are not on the inference path")


class ConvModule(nn.Module):
    """mmcv/cnn/bricks/conv_module.py (1.6.0): order ('conv', 'norm', 'act'), bias='auto' -> bias iff no norm."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias="auto",
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), inplace=True, with_spectral_norm=False,
                 padding_mode="zeros", order=("conv", "norm", "act")):
        super().__init__()
        assert order == ("conv", "norm", "act") and padding_mode == "zeros" and not with_spectral_norm
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                     dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            assert act_cfg["type"] == "ReLU"
            self.activate = nn.ReLU(inplace=inplace)

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.with_norm else None

    def forward(self, x, activate=True, norm=True):
        x = self.conv(x)
        if norm and self.with_norm:
            x = self.norm(x)
        if activate and self.with_activation:
            x = self.activate(x)
        return x


def install():
    """Put the shim into sys.modules as `mmcv` (+ the submodules the reference imports from)."""
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "_peanut_shim", False):
        return
    names = ["mmcv", "mmcv.cnn", "mmcv.cnn.bricks", "mmcv.cnn.bricks.registry", "mmcv.cnn.bricks.transformer",
             "mmcv.cnn.bricks.drop", "mmcv.cnn.utils", "mmcv.cnn.utils.weight_init", "mmcv.runner", "mmcv.utils",
             "mmcv.utils.parrots_wrapper", "mmcv.ops", "mmcv.parallel", "mmcv.runner.base_module"]
    mods = {n: _Lenient(n) for n in names}
    for n, m in mods.items():
        m.__path__ = []
        sys.modules[n] = m
        if "." in n:
            setattr(mods[n.rsplit(".", 1)[0]], n.rsplit(".", 1)[1], m)
    mmcv = mods["mmcv"]
    mmcv._peanut_shim = True
    mmcv.__version__ = "1.6.0"
    models = Registry("model")
    cnn = mods["mmcv.cnn"]
    cnn.MODELS = models
    cnn.ConvModule, cnn.build_conv_layer, cnn.build_norm_layer = ConvModule, build_conv_layer, build_norm_layer
    cnn.build_plugin_layer = build_plugin_layer
    mods["mmcv.cnn.bricks.registry"].ATTENTION = Registry("attention")
    run = mods["mmcv.runner"]
    run.BaseModule, run.Sequential, run.ModuleList = BaseModule, Sequential, ModuleList
    run.auto_fp16, run.force_fp32 = _identity_decorator, _identity_decorator
    mods["mmcv.runner.base_module"].BaseModule = BaseModule
    mods["mmcv.utils"].Registry = Registry
    mods["mmcv.utils.parrots_wrapper"]._BatchNorm = _BatchNorm
