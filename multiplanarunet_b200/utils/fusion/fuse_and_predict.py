"""Multi-view prediction helpers, mirroring mpunet/utils/fusion/{fuse_and_predict,fusion_training}.py.

`predict_volume` / `map_real_space_pred` / `predict_and_map` keep the reference's numpy signatures
(fuse_and_predict.py:81-137, fusion_training.py:40-89); `predict_multi_view` is the fused device path
`mp predict` uses: sampler -> U-Net -> (map + fuse + argmax) without leaving HBM, replacing the loop
of bin/predict.py:294-366.
"""
import ctypes

import numpy as np

from ... import _C
from ..._C import lib, check


def predict_volume(model, X, batch_size=8, axis=0):
    """fuse_and_predict.py:81-89."""
    X = np.moveaxis(X, source=axis, destination=0)
    pred = model.predict(X, batch_size=batch_size, verbose=1)
    return np.moveaxis(pred, source=0, destination=axis)


_CENTER_CACHE = {}


def voxel_grid_center(shape3, affine3x3):
    """The centre get_voxel_grid_real_space subtracts (sample_grid.py:117-118): np.mean over all voxels of
    A.(i,j,k), reproduced bit for bit (numpy's pairwise summation, interpolation/voxel_center.py - leaf sums on the
    device, tree on the host; ~1 ms for 256^3).  The closed form A.((n-1)/2) differs in the last bits for affines
    that are not exact in binary, which can flip a nearest-neighbour tie.  Cached per (shape, affine)."""
    from ...interpolation.voxel_center import voxel_grid_center_exact
    A = np.ascontiguousarray(np.asarray(affine3x3, dtype=np.float64)[:3, :3])
    key = (tuple(int(v) for v in shape3), A.tobytes())
    if key not in _CENTER_CACHE:
        if len(_CENTER_CACHE) > 64:
            _CENTER_CACHE.clear()
        _CENTER_CACHE[key] = voxel_grid_center_exact(shape3, A)
    return _CENTER_CACHE[key]


def linspace_params(axis):
    """(start, step, stop) if `axis` is bit for bit what np.linspace(start, stop, n) returns - numpy builds it as
    arange(n) * step + start with step = (stop - start) / (n - 1) and then overwrites the last element with stop
    (numpy/_core/function_base.py) - else None.  The reference's grids are exactly such arrays (in-plane axis:
    sample_grid.py:227-233 via np.linspace in test mode; offsets: isotrophic_live_view_sequence_2d.py:62)."""
    a = np.ascontiguousarray(axis, dtype=np.float64)
    n = a.shape[0]
    if a.ndim != 1 or n < 2:
        return None
    start, stop = float(a[0]), float(a[-1])
    step = (stop - start) / (n - 1)
    if not (step > 0) or not np.isfinite(step):
        return None
    ref = np.arange(0, n, dtype=np.float64) * step + start
    ref[-1] = stop
    return (start, step, stop) if np.array_equal(ref, a) else None


def _map_fuse(pred_tensors, grids, inv_bases, shape3, affine3x3, W=None, b=None, sum_fusion=False,
              want_labels=True, want_probs=False, want_combined=False, force_tables=False):
    import torch
    V = len(pred_tensors)
    dev = pred_tensors[0].device
    n_planes, dim, _, C = pred_tensors[0].shape
    ax = np.asarray(grids[0][0], dtype=np.float64)
    offsets = np.stack([np.asarray(g[2], dtype=np.float64) for g in grids])
    lin_ax = linspace_params(ax)
    lin_off = [linspace_params(o) for o in offsets]
    same_axes = all(np.array_equal(np.asarray(g[0]), ax) and np.array_equal(np.asarray(g[1]), ax) for g in grids)
    import os
    force_tables = force_tables or os.environ.get("MPU_MAPFUSE_TABLES") == "1"  # bring-up knob
    use_lin = (not force_tables and 2 <= C <= 8 and lin_ax is not None and all(l is not None for l in lin_off)
               and same_axes and all(t.is_contiguous() for t in pred_tensors))
    inv_step = [(len(ax) - 1) / (ax[-1] - ax[0])] + [(offsets.shape[1] - 1) / (o[-1] - o[0]) for o in offsets]
    ax_d = torch.from_numpy(ax).to(dev)
    off_d = torch.from_numpy(np.ascontiguousarray(offsets)).to(dev)
    ib_d = torch.from_numpy(np.ascontiguousarray(np.stack(inv_bases).reshape(V, 9))).to(dev)
    X, Y, Z = [int(s) for s in shape3]
    labels = torch.empty(X, Y, Z, dtype=torch.uint8, device=dev) if want_labels else None
    probs = torch.empty(X, Y, Z, C, dtype=torch.float32, device=dev) if want_probs else None
    combined = torch.empty(V, X, Y, Z, C, dtype=torch.float32, device=dev) if want_combined else None
    Wd = bd = None
    if not sum_fusion:
        Wd = torch.as_tensor(np.asarray(W, dtype=np.float32).reshape(V, C)).to(dev).contiguous()
        bd = torch.as_tensor(np.asarray(b, dtype=np.float32).reshape(C)).to(dev).contiguous()
    ptrs = (ctypes.c_void_p * V)(*[t.data_ptr() for t in pred_tensors])
    mean = voxel_grid_center(shape3, affine3x3)
    if use_lin:
        check(lib.mpu_map_fuse_linspace(ptrs, V, C, dim, n_planes,
                                        _C.double_array(np.stack(inv_bases).astype(np.float64).ravel()),
                                        _C.double_array(lin_ax), _C.double_array(np.asarray(lin_off).ravel()),
                                        _C.int_array([X, Y, Z]),
                                        _C.double_array(np.asarray(affine3x3, dtype=np.float64).ravel()),
                                        _C.double_array(mean), _C.ptr(Wd), _C.ptr(bd), int(bool(sum_fusion)),
                                        _C.ptr(labels), _C.ptr(probs), _C.ptr(combined), _C.current_stream()),
              "mpu_map_fuse_linspace")
        return labels, probs, combined
    check(lib.mpu_map_fuse(ptrs, V, C, dim, n_planes, _C.ptr(ib_d), _C.ptr(ax_d), _C.ptr(off_d),
                           _C.double_array(inv_step), _C.int_array([X, Y, Z]),
                           _C.double_array(np.asarray(affine3x3, dtype=np.float64).ravel()),
                           _C.double_array(mean), _C.ptr(Wd), _C.ptr(bd), int(bool(sum_fusion)),
                           _C.ptr(labels), _C.ptr(probs), _C.ptr(combined), _C.current_stream()),
          "mpu_map_fuse")
    return labels, probs, combined


def map_real_space_pred(pred, grid, inv_basis, voxel_grid_real_space=None, method="nearest",
                        shape3=None, affine3x3=None):
    """fuse_and_predict.py:92-137 with the reference's array layouts: pred [dim,dim,n,C] float32 ->
    mapped [X,Y,Z,C] float32.  The voxel grid is regenerated on the device from (shape, affine); pass
    either those or the reference's `voxel_grid_real_space` [3,X,Y,Z] array (identity-spaced grids)."""
    import torch
    if method != "nearest":
        raise NotImplementedError("only method='nearest' (the one the reference uses) is implemented")
    if shape3 is None:
        vg = np.asarray(voxel_grid_real_space)
        shape3 = vg.shape[1:]
        # recover the 3x3 affine from unit steps of the centred grid
        o = vg[:, 0, 0, 0]
        affine3x3 = np.stack([vg[:, 1, 0, 0] - o if shape3[0] > 1 else np.zeros(3),
                              vg[:, 0, 1, 0] - o if shape3[1] > 1 else np.zeros(3),
                              vg[:, 0, 0, 1] - o if shape3[2] > 1 else np.zeros(3)], axis=1)
    p = torch.as_tensor(np.ascontiguousarray(np.moveaxis(np.asarray(pred, dtype=np.float32), 2, 0))).cuda()
    V1 = np.zeros((1, p.shape[-1]), dtype=np.float32)
    _, _, combined = _map_fuse([p], [grid], [np.asarray(inv_basis)], shape3, affine3x3, W=V1 + 1,
                               b=np.zeros(p.shape[-1]), want_labels=False, want_combined=True)
    return combined[0].cpu().numpy()


def predict_stack_device(model, seq, image, view, n_planes="same+20", batch_size=None):
    """Sample one view's plane stack straight into the U-Net input tensor and run inference.
    Returns (probs [n,dim,dim,C] f32 tensor, grid, inv_basis)."""
    import torch
    ptr, cpad, rows = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_longlong()
    check(lib.mpu_unet_input_buffer(model._h, ctypes.byref(ptr), ctypes.byref(cpad), ctypes.byref(rows)))
    dim = seq.sample_dim
    bs = min(batch_size or model.max_batch, model.max_batch)
    from ...interpolation import plane_basis, view_offsets
    offsets = view_offsets(dim, seq.real_space_span, n_planes)
    basis = plane_basis(view, 0.)
    n = len(offsets)
    probs = torch.empty(n, dim, dim, model.n_classes, dtype=torch.float32, device=model.device)
    in_view = _InputView(ptr.value)
    for s in range(0, n, bs):
        e = min(n, s + bs)
        image.interpolator.sample_planes(basis, offsets[s:e], dim, seq.real_space_span,
                                         center=image.scaler_center, scale=image.scaler_scale,
                                         out_padded=in_view, cpad=cpad.value, want_f32=False,
                                         want_labels=False)
        check(lib.mpu_unet_forward(model._h, e - s, 0, _C.ptr(probs[s:e]), _C.current_stream()),
              "mpu_unet_forward")
    hd = seq.real_space_span // 2
    g = np.linspace(-hd, hd, dim)
    return probs, (g, g, offsets), np.linalg.inv(basis)


class _InputView(object):
    """Raw device pointer with the tiny tensor-like surface _C.ptr() needs."""

    def __init__(self, p):
        self._p = p

    def data_ptr(self):
        return self._p


def predict_and_map(model, seq, image, view, batch_size=None, voxel_grid_real_space=None,
                    targets=None, eval_prob=1.0, n_planes="same+20"):
    """fusion_training.py:40-89: one view's mapped softmax volume, flattened to [N_vox, C] float32."""
    probs, grid, inv_basis = predict_stack_device(model, seq, image, view, n_planes, batch_size)
    C = probs.shape[-1]
    _, _, combined = _map_fuse([probs], [grid], [inv_basis], image.shape[:3], image.affine[:3, :3],
                               W=np.ones((1, C), np.float32), b=np.zeros(C, np.float32),
                               want_labels=False, want_combined=True)
    return combined[0].reshape(-1, C)


def predict_multi_view(model, seq, image, views, fusion_W=None, fusion_b=None, sum_fusion=False,
                       n_planes="same+20", batch_size=None, want_probs=False, want_combined=False):
    """The whole `mp predict` inner loop for one image (bin/predict.py:294-366) on the device.
    Returns (labels uint8 [X,Y,Z] tensor, probs | None, combined | None)."""
    preds, grids, inv_bases = [], [], []
    for v in views:
        p, g, ib = predict_stack_device(model, seq, image, v, n_planes, batch_size)
        preds.append(p)
        grids.append(g)
        inv_bases.append(ib)
    return _map_fuse(preds, grids, inv_bases, image.shape[:3], image.affine[:3, :3], fusion_W, fusion_b,
                     sum_fusion, True, want_probs, want_combined)


def stack_collections(points_collection, targets_collection):
    """fusion_training.py:7-37: concatenate per-image point sets."""
    X = np.concatenate([np.asarray(p) for p in points_collection], axis=0)
    y = np.concatenate([np.asarray(t).reshape(-1, 1) for t in targets_collection], axis=0)
    return X, y
