"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement in numpy of the reference's multi-view mapping + fusion:
  * voxel grid in real space ...... mpunet/interpolation/sample_grid.py:101-130
  * map_real_space_pred ........... mpunet/utils/fusion/fuse_and_predict.py:92-137 (nearest gather,
                                    out-of-bounds -> one-hot background)
  * FusionLayer.call .............. mpunet/models/fusion_model.py:38-39  softmax(sum_v W*x + b)
  * merge + argmax ................ mpunet/bin/predict.py:349-366, mpunet/utils/utils.py:311-328
  * generalized dice loss ......... mpunet/evaluate/loss_functions.py:207-246 (rank-2 inputs)
  * weight regulariser ............ mpunet/models/fusion_model.py:9-11
  * Adam .......................... tf.keras.optimizers.Adam (TF 2.3.2; third party, restated from its
                                    documented update rule - parity unpinned for the optimizer)
  * dice_all ...................... mpunet/evaluate/metrics.py:26-52

map_real_space_pred is pinned against the unmodified reference source under oracle/ref_shim.py
(tests/test_oracle_vs_reference.py, tests/golden/).  The fusion layer, its regulariser and the generalized dice
loss are pinned against the reference's OWN FusionModel / FusionLayer.call / reg / sparse_generalized_dice_loss
executed unmodified under oracle/keras_shim.py, which supplies eager numpy versions of the elementary TensorFlow
ops those few lines are written in (tests/golden/fusion_ref.npz, incl. central-difference gradients of the
reference objective for the analytic gradients below).  The Adam rule stays "parity unpinned" (it lives in
TensorFlow, not in the reference tree).
"""
import numpy as np

from .sampler import find_indices


def voxel_grid_real_space(shape3, affine3x3):
    """sample_grid.py:101-130 -> [3, X, Y, Z] float64 (centred)."""
    grid = np.mgrid[0:shape3[0]:1, 0:shape3[1]:1, 0:shape3[2]:1]
    pts = np.empty((int(np.prod(shape3)), 3), dtype=grid.dtype)
    for i in range(3):
        pts[:, i] = grid[i].ravel()
    real = np.asarray(affine3x3).dot(pts.T).T
    real = real - np.mean(real, axis=0)
    out = np.empty((3,) + tuple(shape3), dtype=real.dtype)
    for i in range(3):
        out[i] = real[:, i].reshape(shape3)
    return out


def map_real_space_pred(pred, grid, inv_basis, vgrid):
    """fuse_and_predict.py:92-137.  pred [dim,dim,n,C] f32; grid=(ax,ax,offsets) float64;
    vgrid [3,X,Y,Z] float64 -> mapped [X,Y,Z,C] f32."""
    C = pred.shape[-1]
    fill = np.zeros(C, dtype=np.float32)
    fill[0] = 1.0
    shp = vgrid.shape[1:]
    pts = np.empty((int(np.prod(shp)), 3), dtype=vgrid.dtype)
    for i in range(3):
        pts[:, i] = vgrid[i].ravel()
    q = np.asarray(inv_basis).dot(pts.T).T
    sel = []
    oob = np.zeros(q.shape[0], dtype=bool)
    for k in range(3):
        i, t, o = find_indices(np.asarray(grid[k]), q[:, k])
        sel.append(np.where(t <= .5, i, i + 1))
        oob |= o
    res = pred[tuple(sel)].copy()
    res[oob] = fill
    return res.reshape(tuple(shp) + (C,))


def softmax_f32(z):
    z = z.astype(np.float32)
    m = z.max(axis=-1, keepdims=True)
    e = np.exp(z - m, dtype=np.float32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)


def fusion_logits(x, W, b):
    """x [N,V,C] f32, W [V,C], b [1,C] or [C] -> z [N,C] f32; views summed in order v=0..V-1."""
    x = np.asarray(x, dtype=np.float32)
    W = np.asarray(W, dtype=np.float32)
    z = np.zeros((x.shape[0], x.shape[2]), dtype=np.float32)
    for v in range(x.shape[1]):
        z = z + W[v][None, :] * x[:, v, :]
    return z + np.asarray(b, dtype=np.float32).reshape(1, -1)


def fusion_forward(x, W, b):
    """FusionLayer.call (fusion_model.py:38-39): softmax over classes of the weighted view sum."""
    return softmax_f32(fusion_logits(x, W, b))


def merge_views(combined, W=None, b=None, sum_fusion=False):
    """bin/predict.py:349-366. combined [V,X,Y,Z,C] f32 -> (probs [X,Y,Z,C], labels uint8)."""
    V = combined.shape[0]
    shp = combined.shape[1:4]
    C = combined.shape[-1]
    if sum_fusion:
        probs = np.sum(combined, axis=0)
    else:
        x = np.moveaxis(combined, 0, -2).reshape(-1, V, C)
        probs = fusion_forward(x, W, b).reshape(tuple(shp) + (C,))
    labels = probs.argmax(-1).astype(np.uint8)
    return probs, labels


def gdl_loss_and_grads(x, y, W, b, reg=1e-6):
    """Fusion training objective on a batch of points (type_weight='uniform', the CLI default
    bin/train_fusion.py:78).  loss_functions.py:207-246 with rank-2 predictions: no spatial reduction,
    dice_c = 2*onehot_c*p_c / (p_c + onehot_c + 1e-6); loss = mean_n(1 - mean_c dice) + reg terms
    (fusion_model.py:9-11).  Returns (loss, dW, db) in float64 (analytic gradients)."""
    x = np.asarray(x, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    bb = np.asarray(b, dtype=np.float64).reshape(-1)
    N, V, C = x.shape
    z = (W[None] * x).sum(1) + bb[None]
    z = z - z.max(-1, keepdims=True)
    e = np.exp(z)
    p = e / e.sum(-1, keepdims=True)
    onehot = np.zeros((N, C))
    onehot[np.arange(N), np.asarray(y).reshape(-1).astype(np.int64)] = 1.0
    eps = 1e-6
    num = 2.0 * onehot * p
    den = p + onehot + eps
    dice = num / den
    loss_pts = 1.0 - dice.mean(-1)
    loss = loss_pts.mean() + reg * (W ** 2).mean() + reg * (bb ** 2).mean()
    # d loss / d p
    ddice_dp = (2.0 * onehot * den - num) / den ** 2
    dp = -ddice_dp / C / N
    # softmax backward
    dz = p * (dp - (dp * p).sum(-1, keepdims=True))
    dW = (dz[:, None, :] * x).sum(0) + reg * 2.0 * W / W.size
    db = dz.sum(0) + reg * 2.0 * bb / bb.size
    return loss, dW, db


def adam_step(theta, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-7):
    """Keras Adam (TF 2.3.2 optimizer_v2/adam.py, non-amsgrad): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    theta -= lr_t * m / (sqrt(v) + eps).  Keras' default epsilon is 1e-7; the reference passes none for
    the fusion model (bin/train_fusion.py:345) and 1e-8 for the U-Net (train_hparams.yaml:125-126)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v


def dice_all(y_true, y_pred, n_classes, ignore_zero=True, smooth=1.0):
    """evaluate/metrics.py:26-52."""
    start = 1 if ignore_zero else 0
    out = np.empty(n_classes - start, dtype=np.float32)
    out.fill(np.nan)
    for c in range(start, n_classes):
        s1 = (y_true == c)
        s2 = (y_pred == c)
        if np.any(s1) or np.any(s2):
            inter = np.logical_and(s1, s2).sum()
            out[c - start] = (smooth + 2 * inter) / (smooth + s1.sum() + s2.sum())
    return out
