"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/mpunet_b200.h
declares; compute entries refuse to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mpunet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpu_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libmpunet_b200.so does not export %s" % s


def test_version_and_error_string(lib):
    assert lib.mpu_version() >= 100
    lib.mpu_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.mpu_last_error(), bytes)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multiplanarunet_b200.models import UNet, FusionModel
    with pytest.raises(RuntimeError):
        UNet(n_classes=3, dim=32)
    with pytest.raises(RuntimeError):
        FusionModel(6, 5)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may touch oracle/."""
    pkg = os.path.join(ROOT, "multiplanarunet_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)
