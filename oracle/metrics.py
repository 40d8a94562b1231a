"""ORACLE (test infrastructure, not the product): numpy restatement of the reference's evaluation counts.

  * cm_counts ........ mpunet/callbacks/validation.py:117-131 (arg-max, then three np.bincount calls)
  * compute_dice ..... mpunet/callbacks/validation.py:60-90
  * dice / dice_all .. mpunet/evaluate/metrics.py:12-52 (dice_all also lives in oracle/fusion.py)

Pinned against the unmodified reference's `dice_all` in tests/test_oracle_vs_reference.py (validation.py needs a
real TensorFlow to import; its three-line counting rule is restated from the source and checked against
`dice_all`-style brute force in tests/test_oracle_golden.py).
"""
import numpy as np


def cm_counts(y_true, pred, n_classes):
    """pred: labels or scores [..., n_classes]; -> (tps, rel, sel) uint64 arrays of length n_classes."""
    p = np.asarray(pred)
    if p.dtype.kind == "f":
        p = p.reshape(-1, n_classes).argmax(-1)
    p = p.ravel().astype(np.int64)
    y = np.asarray(y_true).ravel().astype(np.int64)
    tps = np.bincount(np.where(y == p, y, n_classes), minlength=n_classes + 1)[:-1]
    rel = np.bincount(y, minlength=n_classes)
    sel = np.bincount(p, minlength=n_classes)
    return tps.astype(np.uint64), rel.astype(np.uint64), sel.astype(np.uint64)


def compute_dice(tp, rel, sel):
    sel_mask = sel > 0
    rel_mask = rel > 0
    precisions = np.zeros(shape=tp.shape, dtype=np.float32)
    recalls = np.zeros_like(precisions)
    dices = np.zeros_like(precisions)
    precisions[sel_mask] = tp[sel_mask] / sel[sel_mask]
    recalls[rel_mask] = tp[rel_mask] / rel[rel_mask]
    intrs = (2 * precisions * recalls)
    union = (precisions + recalls)
    dice_mask = union > 0
    dices[dice_mask] = intrs[dice_mask] / union[dice_mask]
    return precisions, recalls, dices


def dice(y_true, y_pred, smooth=1.0):
    s1 = np.array(y_true).flatten().astype(bool)
    s2 = np.array(y_pred).flatten().astype(bool)
    return (smooth + 2 * np.logical_and(s1, s2).sum()) / (smooth + s1.sum() + s2.sum())


def dice_all(y_true, y_pred, smooth=1.0, n_classes=None, ignore_zero=True, skip_if_no_y=False):
    if n_classes is None:
        classes = np.unique(y_true)
    else:
        classes = np.arange(max(2, n_classes))
    if ignore_zero:
        classes = classes[np.where(classes != 0)]
    out = np.empty(shape=classes.shape, dtype=np.float32)
    out.fill(np.nan)
    for idx, c in enumerate(classes):
        s1 = y_true == c
        if skip_if_no_y and not np.any(s1):
            continue
        s2 = y_pred == c
        if np.any(s1) or np.any(s2):
            out[idx] = dice(s1, s2, smooth=smooth)
    return out
