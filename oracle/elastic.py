"""ORACLE (test infrastructure, not the product): numpy restatement of the reference's Elastic2D augmentation.

  * elastic_transform_2d ... mpunet/augmentation/elastic_deformation.py:6-69
  * Elastic.__call__ ....... mpunet/augmentation/augmenters.py:87-109 (mask draw, alpha/sigma draws, aug weight)
  * 2-D grid interpolation . mpunet/interpolation/regular_grid_interpolator.py:204-223,252-270 on arange grids

The Gaussian smoothing is scipy.ndimage.gaussian_filter (third-party, scipy pinned >=1.x by the reference's
requirements; present in this image) - the product restates its separable float64 correlation on the device.
Pinned against the unmodified reference in tests/test_oracle_vs_reference.py and by tests/golden/elastic.npz.
"""
import itertools

import numpy as np
from scipy.ndimage import gaussian_filter

from .sampler import find_indices


def displacement_fields(noise_dx, noise_dy, alpha, sigma):
    """noise_* = np.random.rand(H, W) * 2 - 1 as drawn by the reference (:44-47)."""
    dx = gaussian_filter(noise_dx, sigma, mode="constant", cval=0.) * alpha
    dy = gaussian_filter(noise_dy, sigma, mode="constant", cval=0.) * alpha
    return dx, dy


def _interp2(values, px, py, fill, nearest):
    H, W = values.shape[:2]
    grids = (np.arange(H), np.arange(W))
    idx, ts = [], []
    oob = np.zeros(px.size, dtype=bool)
    for g, x in zip(grids, (px.ravel(), py.ravel())):
        i, t, o = find_indices(g, x)
        idx.append(i)
        ts.append(t)
        oob |= o
    if nearest:
        sel = [np.where(t <= .5, i, i + 1) for i, t in zip(idx, ts)]
        res = values[tuple(sel)].copy()
    else:
        res = 0.
        for edge in itertools.product(*[[i, i + 1] for i in idx]):
            w = 1.
            for e, i, t in zip(edge, idx, ts):
                w = w * np.where(e == i, 1 - t, t)
            res = res + np.asarray(values[edge]) * w
    res[oob] = fill
    return res


def elastic_transform_2d(image, labels, alpha, sigma, bg_val=0.0, noise=None, rng=np.random):
    """image [H,W(,C)] float32, labels [H,W] or None.  `noise` = (noise_dx, noise_dy) to bypass the RNG."""
    if image.ndim == 2:
        image = np.expand_dims(image, axis=-1)
    shape = image.shape[:2]
    channels = image.shape[-1]
    bg = bg_val if isinstance(bg_val, (list, tuple, np.ndarray)) else [bg_val] * channels
    if noise is None:
        ndx = rng.rand(*shape) * 2 - 1
        ndy = rng.rand(*shape) * 2 - 1
    else:
        ndx, ndy = noise
    dx, dy = displacement_fields(ndx, ndy, alpha, sigma)
    x, y = np.mgrid[0:shape[0], 0:shape[1]]
    px, py = x + dx, y + dy
    out = np.empty(shape=image.shape, dtype=image.dtype)
    for c in range(channels):
        out[..., c] = _interp2(image[..., c], px, py, np.float32(bg[c]), nearest=False).reshape(shape)
    lab = None
    if labels is not None:
        lab = _interp2(labels, px, py, 0, nearest=True).reshape(shape).astype(labels.dtype)
    return out, lab


class Elastic2D(object):
    """augmenters.py:10-126 restated: same RNG call order as the reference."""

    def __init__(self, alpha, sigma, apply_prob, aug_weight=0.33):
        self._alpha, self._sigma, self.apply_prob, self.weight = alpha, sigma, apply_prob, aug_weight

    def _draw(self, v, rng):
        return rng.uniform(v[0], v[1], 1)[0] if isinstance(v, (list, tuple)) else v

    def __call__(self, batch_x, batch_y, bg_values, batch_w=None, rng=np.random):
        mask = rng.rand(len(batch_x)) <= self.apply_prob
        ax, ay = [], []
        for i, (aug, x, y, bg) in enumerate(zip(mask, batch_x, batch_y, bg_values)):
            if aug:
                alpha = self._draw(self._alpha, rng)
                sigma = self._draw(self._sigma, rng)
                x, y = elastic_transform_2d(x, y, alpha, sigma, bg, rng=rng)
                if batch_w is not None:
                    batch_w[i] = self.weight
            ax.append(x)
            ay.append(y)
        return (ax, ay, batch_w) if batch_w is not None else (ax, ay)
