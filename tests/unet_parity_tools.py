"""Helpers for the layer-local parity tests of the U-Net engine (tests/test_gpu_unet_baseline.py).

The device keeps every forward activation and every back-propagated activation gradient in its workspace
(csrc/unet.cu `Level`); `mpu_unet_debug_buffer` exposes them.  These helpers read them out as NHWC float32
arrays in the oracle's naming (oracle/unet.py `cap` points), so that the oracle can be teacher-forced in BOTH
directions and every layer's forward and backward arithmetic compared locally:

    forward residual  : oracle(layer(device inputs))              vs device output
    backward residual : oracle(layer backward(device upstream g)) vs device downstream gradient

Expected agreement: identical bf16 rounding points, fp32 accumulation in a different order - i.e. the two
differ by at most ONE bf16 ulp at a small fraction of the elements (values that sit next to a rounding boundary).
"""
import ctypes

import numpy as np

WHICH = {"a1": 0, "a2": 1, "b": 2, "pooled": 3, "u": 4, "bn1": 5, "c2": 6, "c3": 7, "bn2": 8,
         "gout": 9, "s1": 10, "s2": 11, "dcat": 12, "dzu": 13, "dpool": 14}


def _raw(model, level, which):
    from multiplanarunet_b200._C import lib, check
    ptr, rows, C = ctypes.c_void_p(), ctypes.c_longlong(), ctypes.c_int()
    check(lib.mpu_unet_debug_buffer(model._h, level, WHICH[which], ctypes.byref(ptr), ctypes.byref(rows),
                                    ctypes.byref(C)))
    return ptr.value, C.value


def fetch(model, level, which, B, c_logical):
    """-> (interior [B,H,W,c_logical] float32, sum |border|, sum |padded channels|) of a level buffer."""
    import torch
    ptr, C = _raw(model, level, which)
    H, W = model.img_shape[0] >> level, model.img_shape[1] >> level
    lead = 1
    if which in ("pooled", "dpool"):
        H, W = H // 2, W // 2
    if which == "dzu":
        H, W, lead = H // 2, W // 2, 4
    cw = 2 * C if which == "dcat" else C
    n = lead * B * (H + 2) * (W + 2) * cw
    torch.cuda.synchronize()
    off = ptr - model.workspace.data_ptr()
    arr = model.workspace[off:off + 2 * n].view(torch.bfloat16).float().cpu().numpy()
    arr = arr.reshape(lead * B, H + 2, W + 2, cw)
    bm = np.ones(arr.shape[:3], dtype=bool)
    bm[:, 1:-1, 1:-1] = False
    border = float(np.abs(arr[bm]).sum())
    inner = arr[:, 1:-1, 1:-1, :]
    if which == "dcat":
        padc = float(np.abs(inner[..., c_logical:C]).sum() + np.abs(inner[..., C + c_logical:]).sum())
        return (inner[..., :c_logical].copy(), inner[..., C:C + c_logical].copy()), border, padc
    padc = float(np.abs(inner[..., c_logical:]).sum())
    inner = inner[..., :c_logical]
    if which == "dzu":  # phase-major [a*2+b][B][h][w][C] -> [B][2h][2w][C]
        ph = inner.reshape(2, 2, B, H, W, c_logical)
        full = np.zeros((B, 2 * H, 2 * W, c_logical), np.float32)
        for a in range(2):
            for b in range(2):
                full[:, a::2, b::2, :] = ph[a, b]
        inner = full
    return inner.copy(), border, padc


def collect_forward(model, B, chans, depth=4):
    """All forward activations of the last (train-)forward in the oracle's names; asserts zero borders/pads."""
    force = {}
    for l in range(depth + 1):
        for which in ["a1", "a2", "b"] + (["pooled", "u", "bn1", "c2", "c3", "bn2"] if l < depth else []):
            arr, border, padc = fetch(model, l, which, B, chans[l])
            assert border == 0 and padc == 0, ("forward buffer has a dirty border / padded channel", which, l)
            force["%s_%d" % (which, l)] = arr
    return force


def run_staged_backward(model, x, y, sw, chans, depth=4):
    """train_forward + the three backward stages, reading the activation gradients between stages (the up path
    and the encoder share the s1/s2 scratch of a level).  Returns (loss_sum, force, force_grad)."""
    import torch
    from multiplanarunet_b200 import _C
    B = model._pack(x)
    H, W, _ = model.img_shape
    yd = torch.as_tensor(np.ascontiguousarray(y).reshape(B, H, W).astype(np.uint8)).to(model.device)
    swd = None if sw is None else torch.as_tensor(np.asarray(sw, np.float32)).to(model.device)
    st = _C.current_stream()
    gscale = 1.0 if model.loss_scale_mode == "sum" else 1.0 / (B * H * W)
    _C.check(_C.lib.mpu_unet_train_forward(model._h, B, _C.ptr(yd), _C.ptr(swd), ctypes.c_float(gscale),
                                           _C.ptr(model._loss_dev), _C.ptr(None), st), "train_forward")
    force = collect_forward(model, B, chans, depth)
    fg, dirty = {}, []

    def grab(name, level, which, c):
        arr, border, padc = fetch(model, level, which, B, c)
        if border != 0 or padc != 0:
            dirty.append((name, which, level, border, padc))
        return arr

    _C.check(_C.lib.mpu_unet_backward_stage(model._h, B, 0, st), "backward_stage 0")
    for l in range(depth):
        fg["bn2_%d" % l] = grab("bn2", l, "gout", chans[l])
        fg["c3_%d" % l] = grab("c3", l, "s1", chans[l])
        fg["c2_%d" % l] = grab("c2", l, "s2", chans[l])
        skip, up = grab("cat", l, "dcat", chans[l])
        fg["skip_%d" % l], fg["bn1_%d" % l] = skip, up
        fg["u_%d" % l] = grab("u", l, "dzu", chans[l])
    fg["b_%d" % depth] = grab("b", depth, "gout", chans[depth])
    _C.check(_C.lib.mpu_unet_backward_stage(model._h, B, 1, st), "backward_stage 1")
    _C.check(_C.lib.mpu_unet_backward_stage(model._h, B, 2, st), "backward_stage 2")
    for l in range(depth + 1):
        fg["a2_%d" % l] = grab("a2", l, "s1", chans[l])
        fg["a1_%d" % l] = grab("a1", l, "s2", chans[l])
        if l < depth:
            fg["pooled_%d" % l] = grab("pooled", l, "dpool", chans[l])
    torch.cuda.synchronize()
    assert not dirty, ("gradient buffers with non-zero border / padded channels", dirty)
    return float(model._loss_dev.item()), force, fg


RELU_POINTS = ("a1", "a2", "c2", "c3", "u")  # activations stored post-ReLU: the device keeps mask * gradient


def bf16_ulp(v):
    """Spacing of bf16 (8 significant bits) at |v|."""
    a = np.maximum(np.abs(v).astype(np.float64), 1e-30)
    return np.exp2(np.floor(np.log2(a)) - 7.0)


def residual(ref, got, floor_rel=1e-4):
    """-> (max error in units of [bf16 ulp at the larger magnitude + floor], fraction of elements that differ).
    `floor` = floor_rel x rms(ref) absorbs fp32 accumulation noise on values that cancel to ~0."""
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    rms = float(np.sqrt((ref ** 2).mean())) + 1e-30
    tol = bf16_ulp(np.maximum(np.abs(ref), np.abs(got))) + floor_rel * rms
    d = np.abs(ref - got)
    return float((d / tol).max()), float((d > 0).mean())


def worst_elements(ref, got, k=6, floor_rel=1e-4):
    """Diagnostics for a failing residual: the k worst elements (index, ref, got, error in tolerance units) and
    how the out-of-tolerance elements distribute over images / rows / columns / 64-channel groups."""
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    rms = float(np.sqrt((ref ** 2).mean())) + 1e-30
    tol = bf16_ulp(np.maximum(np.abs(ref), np.abs(got))) + floor_rel * rms
    r = np.abs(ref - got) / tol
    bad = r > 1.0
    lines = ["   out of tolerance: %d of %d elements; rms(ref) %.4g" % (int(bad.sum()), bad.size, rms)]
    order = np.argsort(r.ravel())[::-1][:k]
    for o in order:
        idx = np.unravel_index(o, r.shape)
        lines.append("   %s ref %.6g got %.6g (%.1f tol)" % (tuple(int(i) for i in idx), ref[idx], got[idx], r[idx]))
    if bad.any() and bad.ndim == 4:
        nz = np.argwhere(bad)
        for ax, name in ((0, "image"), (1, "y"), (2, "x")):
            vals, cnt = np.unique(nz[:, ax], return_counts=True)
            top = np.argsort(cnt)[::-1][:8]
            lines.append("   by %s: %s" % (name, ", ".join("%d:%d" % (vals[t], cnt[t]) for t in top)))
        vals, cnt = np.unique(nz[:, 3] // 8, return_counts=True)
        top = np.argsort(cnt)[::-1][:12]
        lines.append("   by channel//8: %s" % ", ".join("%d:%d" % (vals[t], cnt[t]) for t in top))
    return "\n".join(lines)
