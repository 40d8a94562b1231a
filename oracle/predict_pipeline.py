"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the reference's multi-view prediction loop for one image, with the reference's own threading:
  * _multi_view_predict_on .......... mpunet/bin/predict.py:294-346
  * get_view_from (7 worker threads) mpunet/sequences/isotrophic_live_view_sequence_2d.py:29-101
  * predict_volume ................... mpunet/utils/fusion/fuse_and_predict.py:81-89 (batch_size = 8 slices)
  * map_real_space_pred (7 threads
    over x-slabs) .................... mpunet/utils/fusion/fuse_and_predict.py:92-137
  * get_voxel_grid_real_space ........ mpunet/interpolation/sample_grid.py:101-130
  * merge_multi_view_preds ........... mpunet/bin/predict.py:349-366
built from the pinned pieces of oracle/sampler.py and oracle/fusion.py (each checked against the unmodified reference,
tests/test_oracle_golden.py, tests/test_oracle_vs_reference.py).  The U-Net forward is passed in as a callable
(oracle/unet.py for timing; in the path-level parity test the device's own per-view probabilities are fed instead, so
that the comparison isolates sampler -> mapping -> fusion -> argmax).

Used by tests/test_gpu_predict_path.py (label-map parity of `mp predict`) and by bench.py's cpu_baseline /
`--impl reference` legs, where `plane_subset` / `slab_subset` bound the CPU work and the per-unit times are scaled to
the whole volume (the sample is reported).
"""
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import fusion, sampler


def get_view_from_threaded(vol, labels, affine, view, dim, span, bg_value, center, scale, n_planes="same+20",
                           max_workers=7, plane_subset=None):
    """sampler.get_view_from with the reference's ThreadPoolExecutor(max_workers=7) over planes.
    plane_subset: indices of the planes to actually sample (timing mode; the others stay zero)."""
    pix = np.linalg.norm(affine[:3, :3], axis=0)
    rot = None
    if np.any(~np.isclose(np.diag(pix), affine[:3, :3])):
        rot = np.diag(pix).dot(np.linalg.inv(affine[:3, :3]))
    basis = sampler.plane_basis(view)
    offsets = sampler.view_offsets(dim, span, n_planes)
    n = len(offsets)
    X = np.zeros((dim, dim, n, vol.shape[-1]), dtype=np.float32)
    y = np.zeros((dim, dim, n), dtype=np.uint8) if labels is not None else None
    todo = range(n) if plane_subset is None else plane_subset

    def _do(k):
        return k, sampler.sample_plane(vol, labels, pix, basis, dim, span, offsets[k], bg_value, 0, center, scale,
                                       rot_mat=rot)

    with ThreadPoolExecutor(max_workers=max_workers) as pool:
        for k, (im, lab) in pool.map(_do, todo):
            X[:, :, k, :] = im
            if y is not None:
                y[:, :, k] = lab
    hd = span // 2
    g = np.linspace(-hd, hd, dim)
    return X, y, (g, g, offsets), np.linalg.inv(basis)


def map_real_space_pred_threaded(pred, grid, inv_basis, vgrid, max_workers=7, slab_subset=None):
    """fusion.map_real_space_pred with the reference's per-x-slab thread pool (fuse_and_predict.py:118-134)."""
    C = pred.shape[-1]
    fill = np.zeros(C, dtype=np.float32)
    fill[0] = 1.0
    shp = vgrid.shape[1:]
    pts = np.empty((int(np.prod(shp)), 3), dtype=vgrid.dtype)
    for i in range(3):
        pts[:, i] = vgrid[i].ravel()
    q = np.asarray(inv_basis).dot(pts.T).T.reshape(tuple(shp) + (3,))
    mapped = np.zeros(tuple(shp) + (C,), dtype=pred.dtype)
    todo = range(shp[0]) if slab_subset is None else slab_subset

    def _do(ix):
        qs = q[ix].reshape(-1, 3)
        sel = []
        oob = np.zeros(qs.shape[0], dtype=bool)
        for k in range(3):
            i, t, o = sampler.find_indices(np.asarray(grid[k]), qs[:, k])
            sel.append(np.where(t <= .5, i, i + 1))
            oob |= o
        res = pred[tuple(sel)].copy()
        res[oob] = fill
        return ix, res.reshape(shp[1], shp[2], C)

    with ThreadPoolExecutor(max_workers=max_workers) as pool:
        for ix, res in pool.map(_do, todo):
            mapped[ix] = res
    return mapped


def predict_multi_view(vol, affine, views, dim, span, bg_value, center, scale, unet_predict=None, W=None, b=None,
                       sum_fusion=False, n_planes="same+20", per_view_probs=None, labels=None):
    """The whole loop for one image.  unet_predict: callable [n,dim,dim,C] float32 -> [n,dim,dim,K] float32, or
    per_view_probs: list of [n,dim,dim,K] arrays used instead.  Returns (label map uint8, probs, combined, X stacks)."""
    vgrid = fusion.voxel_grid_real_space(vol.shape[:3], affine[:3, :3])
    combined, stacks = [], []
    for v, view in enumerate(views):
        X, _, grid, inv_basis = get_view_from_threaded(vol, labels, affine, view, dim, span, bg_value, center, scale,
                                                       n_planes)
        stacks.append(X)
        if per_view_probs is not None:
            pred = np.moveaxis(np.asarray(per_view_probs[v], dtype=np.float32), 0, 2)
        else:
            pred = np.moveaxis(unet_predict(np.moveaxis(X, 2, 0)), 0, 2)
        combined.append(map_real_space_pred_threaded(pred, grid, inv_basis, vgrid))
    combined = np.stack(combined)
    probs, label_map = fusion.merge_views(combined, W, b, sum_fusion)
    return label_map, probs, combined, stacks


def time_predict_sample(vol, affine, views, dim, span, bg_value, center, scale, unet_predict, W, b, n_planes="same+20",
                        planes_sampled=14, slices_forward=8, slabs_mapped=16, fuse_fraction=1 / 16.0, n_classes=5):
    """Bounded timing of the reference-shaped CPU pipeline: every stage runs on a sample and is scaled to the whole
    volume (stages are embarrassingly parallel over their units: planes, slices, x-slabs, voxels).
    Returns {stage: seconds per VOLUME (all views)}, the total, and a description of the sample."""
    V = len(views)
    n = len(sampler.view_offsets(dim, span, n_planes))
    shape = vol.shape[:3]
    out = {}
    idx = np.linspace(0, n - 1, planes_sampled).astype(int)
    t0 = time.perf_counter()
    X, _, grid, inv_basis = get_view_from_threaded(vol, None, affine, views[0], dim, span, bg_value, center, scale,
                                                   n_planes, plane_subset=list(idx))
    out["get_view_from"] = (time.perf_counter() - t0) / planes_sampled * n * V
    xb = np.ascontiguousarray(np.moveaxis(X[:, :, idx[:slices_forward], :], 2, 0))
    unet_predict(xb[:1])  # warm-up (primitive creation)
    t0 = time.perf_counter()
    unet_predict(xb)
    out["unet_forward"] = (time.perf_counter() - t0) / slices_forward * n * V
    sub = (min(shape[0], 64), shape[1], shape[2])
    t0 = time.perf_counter()
    vg_sub = fusion.voxel_grid_real_space(sub, affine[:3, :3])
    out["voxel_grid"] = (time.perf_counter() - t0) * shape[0] / sub[0]
    vgrid = vg_sub - (np.asarray(affine[:3, :3]).dot((np.asarray(shape) - np.asarray(sub)) / 2.0))[:, None, None, None]
    pred = np.random.RandomState(0).rand(dim, dim, n, n_classes).astype(np.float32)
    slabs = list(range(0, sub[0], max(1, sub[0] // slabs_mapped)))[:slabs_mapped]
    t0 = time.perf_counter()
    mapped = map_real_space_pred_threaded(pred, grid, inv_basis, vgrid, slab_subset=slabs)
    out["map_real_space_pred"] = (time.perf_counter() - t0) / len(slabs) * shape[0] * V
    nvox = int(np.prod(shape))
    m = max(1, int(nvox * fuse_fraction))
    xs = np.random.RandomState(1).rand(m, V, n_classes).astype(np.float32)
    t0 = time.perf_counter()
    p = fusion.fusion_forward(xs, W, b)
    p.argmax(-1).astype(np.uint8)
    out["fusion_argmax"] = (time.perf_counter() - t0) / m * nvox
    total = float(sum(out.values()))
    sample = ("%d of %d planes of 1 view (7 threads), U-Net forward of %d slices, voxel grid of a %dx%dx%d slab, mapping "
              "of %d x-slabs of 1 view (7 threads), fusion of 1/%d of the voxels; each scaled to %d views x %d planes / "
              "%d voxels" % (planes_sampled, n, slices_forward, sub[0], sub[1], sub[2], len(slabs),
                             int(round(1 / fuse_fraction)), V, n, nvox))
    del mapped
    return out, total, sample
