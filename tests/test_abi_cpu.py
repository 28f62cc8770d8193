"""CPU-side checks of the boundary: the shared library loads, exports every symbol the header declares,
and refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

from peanut_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "peanut_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = _lib.load()
    declared = _header_symbols()
    assert declared, "no symbols parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/peanut_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == declared, "ctypes prototypes out of sync with the header"
    assert lib.pn_abi_version() >= 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        _lib.Context(0)


def test_product_does_not_import_oracle():
    """The product package must never route through the oracle (tier rule ③)."""
    pkg = os.path.join(ROOT, "peanut_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_precision_codes_match_the_header():
    """The three arithmetic modes of include/peanut_b200.h's pn_precision enum, by the names the shims accept; anything else is
    rejected instead of being mapped to a neighbouring mode."""
    text = open(os.path.join(ROOT, "include", "peanut_b200.h")).read()
    for name, code in (("bf16", "PN_BF16"), ("tf32", "PN_TF32"), ("fp32", "PN_FP32")):
        m = re.search(rf"\b{code}\s*=\s*(\d+)", text)
        assert m and _lib.precision_code(name) == int(m.group(1)) == getattr(_lib, code)
    for bad in ("fp16", "float32", "", None):
        with pytest.raises(ValueError):
            _lib.precision_code(bad)
    assert _lib.load().pn_abi_version() == 3


def test_device_map_dataset_paths_never_fall_back(tmp_path):
    """N4: host arrays take the reference's numpy expression; a CUDA tensor can only go through the kernels (no GPU here, so
    building a device sequence must raise rather than quietly staying on the host)."""
    import numpy as np
    import torch
    from peanut_b200 import map_dataset as D
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    q = D.quantize_full_map(np.full((2, 3, 3), 0.5, np.float32))
    assert q.dtype == np.uint8 and int(q[0, 0, 0]) == 127
    with pytest.raises((RuntimeError, AssertionError)):
        D.DeviceMapSequence(np.zeros((2, 14, 8, 8), np.uint8))
    with pytest.raises(TypeError):
        D.quantize_full_map_device(torch.zeros(4))
