"""Evaluation metrics of mpunet on the device (mpunet/evaluate/metrics.py:12-52,
mpunet/callbacks/validation.py:60-131).

The label volumes stay in HBM: one pass of `mpu_label_counts` produces, per class, the three integer sums every
metric here is built from (|true==c & pred==c|, |true==c|, |pred==c|); the few floating-point operations on
those 3 x n_classes integers follow the reference expressions on the host.
"""
import ctypes

import numpy as np

from .. import _C
from .._C import check, lib


def _dev_u8(a, device):
    import torch
    if torch.is_tensor(a):
        t = a
    else:
        arr = np.asarray(a)
        if arr.dtype != np.uint8:
            if arr.size and (arr.min() < 0 or arr.max() > 255):
                raise ValueError("label values outside [0, 255] are not supported on the device path")
            arr = arr.astype(np.uint8)
        t = torch.from_numpy(np.ascontiguousarray(arr))
    return t.to(device=device, dtype=torch.uint8).contiguous().reshape(-1)


def label_counts(y_true, y_pred, n_classes, counts=None, device=None):
    """-> int64 device tensor [3, n_classes] (TP, relevant, selected), accumulated into `counts` if given.

    y_pred: label array (any integer dtype in [0, 255]) or per-class scores [..., n_classes] float32, in which
    case the first arg-max is the label (callbacks/validation.py:123)."""
    import torch
    _C.require_cuda()
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    yt = _dev_u8(y_true, device)
    n = yt.numel()
    scores = None
    yp = None
    is_scores = (torch.is_tensor(y_pred) and y_pred.is_floating_point()) or (
        not torch.is_tensor(y_pred) and np.asarray(y_pred).dtype.kind == "f")
    if is_scores:
        scores = y_pred if torch.is_tensor(y_pred) else torch.from_numpy(np.ascontiguousarray(y_pred))
        scores = scores.to(device=device, dtype=torch.float32).contiguous()
        if scores.numel() != n * n_classes:
            raise ValueError("scores must hold n_classes values per label (%d != %d x %d)"
                             % (scores.numel(), n, n_classes))
    else:
        yp = _dev_u8(y_pred, device)
        if yp.numel() != n:
            raise ValueError("y_true and y_pred differ in size (%d vs %d)" % (n, yp.numel()))
    if counts is None:
        counts = torch.zeros(3, n_classes, dtype=torch.int64, device=device)
    check(lib.mpu_label_counts(_C.ptr(yt), _C.ptr(yp), _C.ptr(scores), ctypes.c_longlong(n), int(n_classes),
                               _C.ptr(counts), _C.current_stream()), "mpu_label_counts")
    return counts


def dice(y_true, y_pred, smooth=1.0):
    """Soerensen dice of two binary sets (evaluate/metrics.py:12-23)."""
    import torch

    def as_bool(a):
        return (a != 0) if torch.is_tensor(a) else (np.asarray(a) != 0)
    c = label_counts(as_bool(y_true), as_bool(y_pred), 2).cpu().numpy()
    return (smooth + 2 * int(c[0, 1])) / (smooth + int(c[1, 1]) + int(c[2, 1]))


def dice_all(y_true, y_pred, smooth=1.0, n_classes=None, ignore_zero=True, skip_if_no_y=False):
    """Per-class dice (evaluate/metrics.py:26-52): float32 array over the evaluated classes, NaN where a class
    occurs in neither volume (or not in y_true with skip_if_no_y)."""
    k = 256 if n_classes is None else max(2, int(n_classes))
    c = label_counts(y_true, y_pred, k).cpu().numpy()
    if n_classes is None:
        classes = np.nonzero(c[1])[0]  # np.unique(y_true)
    else:
        classes = np.arange(k)
    if ignore_zero:
        classes = classes[classes != 0]
    out = np.empty(shape=classes.shape, dtype=np.float32)
    out.fill(np.nan)
    for idx, cls in enumerate(classes):
        tp, rel, sel = int(c[0, cls]), int(c[1, cls]), int(c[2, cls])
        if skip_if_no_y and rel == 0:
            continue
        if rel or sel:
            out[idx] = (smooth + 2 * tp) / (smooth + rel + sel)
    return out


def compute_dice(tp, rel, sel):
    """Per-class precision tp / sel, recall tp / rel and their harmonic mean (dice) from accumulated confusion counts;
    a class without selected (relevant) pixels gets precision (recall) 0, and dice 0 when both vanish.  float32 results,
    bit-equal to `Validation._compute_dice` (callbacks/validation.py:60-90; pinned in tests/test_oracle_vs_reference.py)."""
    tp, rel, sel = (np.asarray(a, dtype=np.float64) for a in (tp, rel, sel))
    precision = np.divide(tp, sel, out=np.zeros_like(tp), where=sel > 0).astype(np.float32)
    recall = np.divide(tp, rel, out=np.zeros_like(tp), where=rel > 0).astype(np.float32)
    both = precision + recall
    dice = np.divide(2 * precision * recall, both, out=np.zeros_like(both), where=both > 0)
    return precision, recall, dice
