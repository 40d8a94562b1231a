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


def test_wgrad_plan_invariants():
    """The shape-only work decomposition of the weight-gradient GEMM (wave-aware split-K sizing, partial ci tile, launch
    order table) through mpu_debug_wgrad_plan - no device needed.  For the U-Net's own shapes and a sweep of others:
    every (ci tile, K block) is covered exactly once, the grid fills whole waves of 148 CTAs to >= 85 % unless the
    problem is smaller than one wave, the order table is a permutation of the (tile, split) entries sorted by first K
    block, and the partial tile's K ranges are longer than the full tiles'."""
    import ctypes
    import numpy as np
    from multiplanarunet_b200._C import lib
    out = (ctypes.c_int * (12 + 512))()

    def plan(cx, cy, B, H, W):
        wp = W + 2
        taps = (ctypes.c_int * 9)(*[(ky - 1) * wp + (kx - 1) for ky in range(3) for kx in range(3)])
        rows = B * (H + 2) * wp
        assert lib.mpu_debug_wgrad_plan(cx, cy, ctypes.c_longlong(rows), 9, taps, None, out) == 0
        keys = ("grid", "splits", "splits_part", "kbps", "kbps_part", "kblocks", "ci_full", "ci_tiles", "co_tiles",
                "ngroups", "n_entries", "CA")
        d = dict(zip(keys, out[:12]))
        d["order"] = [x & 0xffffffff for x in out[12:12 + d["n_entries"]]]
        d["rows"] = rows
        return d

    shapes = [(96, 96, 32, 256, 256), (96, 184, 32, 128, 128), (184, 184, 32, 128, 128), (184, 368, 32, 64, 64),
              (368, 368, 32, 64, 64), (728, 728, 32, 32, 32), (1448, 1448, 32, 16, 16), (728, 1448, 32, 16, 16),
              (64, 64, 8, 64, 64), (24, 48, 2, 32, 32), (8, 96, 32, 256, 256), (200, 72, 5, 48, 80), (1448, 728, 32, 32, 32)]
    for cx, cy, B, H, W in shapes:
        d = plan(cx, cy, B, H, W)
        assert d["ngroups"] == 3 and d["kblocks"] == (d["rows"] + 63) // 64
        atoms = (cx + 63) // 64
        assert d["CA"] == min(2, atoms) and d["ci_full"] == atoms // d["CA"] and d["co_tiles"] == (cy + 127) // 128
        has_part = atoms % d["CA"] != 0
        assert d["ci_tiles"] == d["ci_full"] + int(has_part)
        # K coverage: splits x range >= all K blocks, and no empty trailing split
        assert d["splits"] * d["kbps"] >= d["kblocks"] > (d["splits"] - 1) * d["kbps"]
        if has_part:
            assert d["splits_part"] * d["kbps_part"] >= d["kblocks"] > (d["splits_part"] - 1) * d["kbps_part"]
            assert d["kbps_part"] >= d["kbps"]            # the lighter CTAs take longer K ranges
        else:
            assert d["splits_part"] == 0
        per = d["ngroups"] * d["co_tiles"]
        assert d["grid"] == per * (d["ci_full"] * d["splits"] + d["splits_part"])
        # whole waves: unless there is less than one wave of work, the last wave is (nearly) full
        units = per * d["ci_tiles"]
        if units * d["kblocks"] >= 148 * 64 and units <= 12 * 148:
            waves = -(-d["grid"] // 148)
            assert d["grid"] >= 0.85 * waves * 148, (cx, cy, H, d["grid"])
        if d["n_entries"]:
            assert d["n_entries"] == d["ci_full"] * d["splits"] + d["splits_part"]
            full = sorted((e >> 16 & 0x7fff, e & 0xffff) for e in d["order"] if not e >> 31)
            part = sorted(e & 0xffff for e in d["order"] if e >> 31)
            assert full == [(c, s) for c in range(d["ci_full"]) for s in range(d["splits"])]
            assert part == list(range(d["splits_part"]))
            starts = [(e & 0xffff) * (d["kbps_part"] if e >> 31 else d["kbps"]) for e in d["order"]]
            assert starts == sorted(starts)
    # the benchmark's own level-0 / level-2 cases: one resp. one wave, no third almost-empty wave (round 1: 297 CTAs)
    assert plan(96, 96, 32, 256, 256)["grid"] == 147 and plan(368, 368, 32, 64, 64)["grid"] in (135, 270)
