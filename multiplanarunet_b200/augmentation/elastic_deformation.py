"""Elastic deformation of 2-D slices on the device (mpunet/augmentation/elastic_deformation.py:6-69).

Host side: the random numbers (drawn with numpy in the reference's order) and the normalised Gaussian taps;
device side (`mpu_elastic_2d`): the separable float64 Gaussian filter of the noise images, the displaced
sampling grid, bilinear image / nearest label resampling with out-of-bounds fill.
"""
import ctypes

import numpy as np

from .. import _C
from .._C import check, lib


def gaussian_taps(sigma, truncate=4.0):
    """Taps of scipy.ndimage.gaussian_filter1d(order=0): radius int(truncate*sigma + 0.5), exp(-x^2/2s^2)
    normalised in float64."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    return phi / phi.sum(), radius


def elastic_batch_device(x, y, noise, alphas, sigmas, bg_values):
    """x [n,H,W,C] f32 cuda, y [n,H,W] u8 cuda or None, noise [n,2,H,W] f64 (host or device),
    alphas/sigmas [n], bg_values [n][C] -> (x_out, y_out) new device tensors."""
    import torch
    _C.require_cuda()
    n, H, W, C = x.shape
    dev = x.device
    taps = [gaussian_taps(s) for s in sigmas]
    wmax = max(2 * r + 1 for _, r in taps)
    wtab = np.zeros((n, wmax), dtype=np.float64)
    for i, (w, r) in enumerate(taps):
        wtab[i, :2 * r + 1] = w
    fields = torch.as_tensor(noise, dtype=torch.float64).to(dev).contiguous().clone()
    scratch = torch.empty_like(fields)
    d_w = torch.from_numpy(wtab).to(dev)
    d_r = torch.tensor([r for _, r in taps], dtype=torch.int32, device=dev)
    d_a = torch.tensor([float(a) for a in alphas], dtype=torch.float64, device=dev)
    d_bg = torch.tensor(np.asarray(bg_values, dtype=np.float32).reshape(n, C), device=dev)
    x = x.contiguous()
    x_out = torch.empty_like(x)
    y_out = None
    if y is not None:
        y = y.contiguous()
        y_out = torch.empty_like(y)
    check(lib.mpu_elastic_2d(_C.ptr(x), _C.ptr(y), _C.ptr(fields), _C.ptr(scratch), _C.ptr(d_w), int(wmax),
                             _C.ptr(d_r), _C.ptr(d_a), _C.ptr(d_bg), int(n), int(H), int(W), int(C),
                             _C.ptr(x_out), _C.ptr(y_out), _C.current_stream()), "mpu_elastic_2d")
    return x_out, y_out


def elastic_transform_2d(image, labels, alpha, sigma, bg_val=0.0, rng=np.random):
    """Same call as the reference function: numpy (or device) image [H,W(,C)] + labels [H,W] -> deformed pair
    of the input kind.  Draws `rand(H, W)` twice from `rng` exactly like the reference."""
    import torch
    is_np = not torch.is_tensor(image)
    im = torch.as_tensor(np.asarray(image, dtype=np.float32)) if is_np else image
    squeeze = im.ndim == 2
    if squeeze:
        im = im[..., None]
    H, W, C = im.shape
    bg = bg_val if isinstance(bg_val, (list, tuple, np.ndarray)) else [bg_val] * C
    noise = np.stack([rng.rand(H, W) * 2 - 1, rng.rand(H, W) * 2 - 1])[None]
    lab = None
    lab_dtype = None
    if labels is not None:
        lab_dtype = None if torch.is_tensor(labels) else np.asarray(labels).dtype
        lab = torch.as_tensor(np.asarray(labels).astype(np.uint8)) if not torch.is_tensor(labels) else labels
        lab = lab.cuda()[None]
    xo, yo = elastic_batch_device(im.cuda().float()[None], lab, noise, [alpha], [sigma], [list(bg)])
    xo = xo[0]
    if is_np:
        xo = xo.cpu().numpy().astype(np.asarray(image).dtype)
    if yo is not None:
        yo = yo[0]
        if lab_dtype is not None:
            yo = yo.cpu().numpy().astype(lab_dtype)
    return xo, yo
