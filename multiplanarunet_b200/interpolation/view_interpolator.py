"""Device-resident counterpart of mpunet.interpolation.view_interpolator.ViewInterpolator.

Holds the image volume (and labels) in HBM once and samples whole stacks / batches of oblique planes
with one kernel launch (csrc/volume_ops.cu: trilinear image, nearest labels, out-of-bounds fill,
optional RobustScaler affine, optional direct write into the U-Net's bf16 input tensor).
Reference: view_interpolator.py:17-147 + regular_grid_interpolator.py:152-270.
"""
import ctypes

import numpy as np

from .. import _C
from .._C import lib, check
from .sample_grid import get_voxel_axes_real_space


class ViewInterpolator(object):
    def __init__(self, image, labels, affine, bg_value=0.0, bg_class=0, device=None, logger=None):
        import torch
        if image.ndim != 4:
            raise ValueError("Input img of dim %i must be dim 4. If image has only 1 channel, use "
                             "np.expand_dims(img, -1)." % image.ndim)
        if not torch.cuda.is_available():
            raise RuntimeError("ViewInterpolator needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.im_shape = tuple(image.shape)
        self.n_channels = self.im_shape[-1]
        self.im_dtype = np.float32
        if not isinstance(bg_value, (list, tuple, np.ndarray)):
            bg_value = [bg_value] * self.n_channels
        if len(bg_value) != self.n_channels:
            raise ValueError("'bg_value' should be a list of length 'n_channels'. Got {} for "
                             "n_channels={}".format(bg_value, self.n_channels))
        self.bg_value = [float(np.float32(b)) for b in bg_value]
        self.bg_class = int(bg_class)
        (gx, gy, gz), self.rot_mat = get_voxel_axes_real_space(self.im_shape[:3], affine)
        if np.any(np.sign(np.diagonal(np.diag(np.linalg.norm(np.asarray(affine)[:3, :3], axis=0)))) == -1):
            raise AssertionError("axes must be strictly increasing")
        self._axes_np = (gx, gy, gz)
        t = torch.as_tensor
        self.vol = (image if torch.is_tensor(image) else t(np.ascontiguousarray(image, dtype=np.float32))).to(
            self.device, dtype=torch.float32).contiguous()
        self.labels = None
        if labels is not None:
            self.labels = (labels if torch.is_tensor(labels) else t(np.ascontiguousarray(labels).astype(np.uint8))).to(
                self.device, dtype=torch.uint8).contiguous()
        self._g = [t(g).to(self.device) for g in (gx, gy, gz)]
        self._inv_step = _C.double_array([
            (len(g) - 1) / (float(g[-1]) - float(g[0])) if len(g) > 1 else 0.0 for g in (gx, gy, gz)])

    def sample_planes(self, bases, offsets, dim, span, center=None, scale=None, out_f32=None,
                      out_padded=None, cpad=0, out_labels=None, want_f32=True, want_labels=True):
        """Sample n planes.  bases [n,3,3] (or one [3,3]) float64, offsets [n] float64.
        Returns (im [n,dim,dim,C] float32 tensor | None, lab [n,dim,dim] uint8 tensor | None)."""
        import torch
        offsets = np.atleast_1d(np.asarray(offsets, dtype=np.float64))
        n = offsets.shape[0]
        bases = np.asarray(bases, dtype=np.float64)
        if bases.ndim == 2:
            bases = np.broadcast_to(bases, (n, 3, 3))
        planes = np.empty((n, 10), dtype=np.float64)
        planes[:, :9] = bases.reshape(n, 9)
        planes[:, 9] = offsets
        planes_d = torch.from_numpy(planes).to(self.device)
        if want_f32 and out_f32 is None:
            out_f32 = torch.empty(n, dim, dim, self.n_channels, dtype=torch.float32, device=self.device)
        if want_labels and self.labels is not None and out_labels is None:
            out_labels = torch.empty(n, dim, dim, dtype=torch.uint8, device=self.device)
        dims = _C.int_array(self.im_shape[:3])
        rot = _C.double_array(np.asarray(self.rot_mat, dtype=np.float64).ravel()) if self.rot_mat is not None else None
        bg = _C.float_array(self.bg_value)
        cen = _C.double_array(center) if center is not None else None
        scl = _C.double_array(scale) if scale is not None else None
        check(lib.mpu_sample_planes(_C.ptr(self.vol), _C.ptr(self.labels), dims, self.n_channels,
                                    _C.ptr(self._g[0]), _C.ptr(self._g[1]), _C.ptr(self._g[2]),
                                    self._inv_step, rot, _C.ptr(planes_d), n, int(dim),
                                    ctypes.c_double(float(span)), bg, self.bg_class, cen, scl,
                                    _C.ptr(out_f32), _C.ptr(out_padded), int(cpad), _C.ptr(out_labels),
                                    _C.current_stream()), "mpu_sample_planes")
        return out_f32, out_labels

    def probe_planes(self, bases, offsets, dim, span):
        """Class-presence bit masks and `is_valid_im` flags of n candidate planes without materialising them
        (mpu_probe_planes).  Returns (class_mask uint32 [n], valid uint32 [n]) device tensors."""
        import torch
        offsets = np.atleast_1d(np.asarray(offsets, dtype=np.float64))
        n = offsets.shape[0]
        planes = np.empty((n, 10), dtype=np.float64)
        planes[:, :9] = np.asarray(bases, dtype=np.float64).reshape(n, 9)
        planes[:, 9] = offsets
        planes_d = torch.from_numpy(planes).to(self.device, non_blocking=True)
        masks = torch.empty(2, n, dtype=torch.int32, device=self.device)
        rot = _C.double_array(np.asarray(self.rot_mat, dtype=np.float64).ravel()) if self.rot_mat is not None else None
        check(lib.mpu_probe_planes(_C.ptr(self.vol), _C.ptr(self.labels), _C.int_array(self.im_shape[:3]),
                                   self.n_channels, _C.ptr(self._g[0]), _C.ptr(self._g[1]), _C.ptr(self._g[2]),
                                   self._inv_step, rot, _C.ptr(planes_d), n, int(dim), ctypes.c_double(float(span)),
                                   _C.float_array(self.bg_value), self.bg_class, _C.ptr(masks[0]), _C.ptr(masks[1]),
                                   _C.current_stream()), "mpu_probe_planes")
        return masks[0], masks[1]

    # -- reference call surface on explicit grids (view_interpolator.py:54-101) --------------------------
    def apply_rotation(self, mgrid):
        """Grid [3, ...] float64 -> grid aligned with the voxel axes (`rot_mat . points`), host numpy like the
        reference (3 x N doubles)."""
        if self.rot_mat is None:
            return mgrid
        from .sample_grid import mgrid_to_points, points_to_mgrid
        shape = mgrid[0].shape
        rotated = np.asarray(self.rot_mat).dot(mgrid_to_points(mgrid).T).T
        return points_to_mgrid(rotated, shape)

    def _interp_points(self, mgrid, want_image, want_labels):
        import torch
        _C.require_cuda()
        grid = np.asarray(mgrid if not isinstance(mgrid, tuple) else np.stack(mgrid), dtype=np.float64)
        if grid.shape[0] != 3:
            raise ValueError("grid must be [3, ...] (x, y, z coordinates); got shape %s" % (grid.shape,))
        out_shape = np.squeeze(grid[0]).shape
        n = int(np.prod(grid.shape[1:]))
        if n == 0:
            return (np.empty(out_shape + (self.n_channels,), np.float32) if want_image else None,
                    np.empty(out_shape, np.uint8) if want_labels else None)
        coords = torch.from_numpy(np.ascontiguousarray(grid.reshape(3, n))).to(self.device)
        im = torch.empty(n, self.n_channels, dtype=torch.float32, device=self.device) if want_image else None
        lab = torch.empty(n, dtype=torch.uint8, device=self.device) if want_labels else None
        check(lib.mpu_interp_points(_C.ptr(self.vol), _C.ptr(self.labels), _C.int_array(self.im_shape[:3]),
                                    self.n_channels, _C.ptr(self._g[0]), _C.ptr(self._g[1]), _C.ptr(self._g[2]),
                                    self._inv_step, _C.ptr(coords), ctypes.c_longlong(n),
                                    _C.float_array(self.bg_value), self.bg_class, _C.ptr(im), _C.ptr(lab),
                                    _C.current_stream()), "mpu_interp_points")
        im_np = im.cpu().numpy().reshape(out_shape + (self.n_channels,)) if want_image else None
        lab_np = lab.cpu().numpy().reshape(out_shape) if want_labels else None
        return im_np, lab_np

    def intrp_image(self, mgrid, apply_rot=True):
        """-> image [h, w, C] float32 at the grid points (trilinear, out-of-bounds = bg_value)."""
        if apply_rot:
            mgrid = self.apply_rotation(mgrid)
        return self._interp_points(mgrid, True, False)[0]

    def intrp_labels(self, mgrid, apply_rot=True):
        """-> labels [h, w] uint8 (nearest, out-of-bounds = bg_class), None without a label volume."""
        if self.labels is None:
            return None
        if apply_rot:
            mgrid = self.apply_rotation(mgrid)
        return self._interp_points(mgrid, False, True)[1]

    def __call__(self, rgrid_mgrid):
        """(image, labels) on one grid, as `image.interpolator(mgrid)` returns them in the reference."""
        grid = self.apply_rotation(rgrid_mgrid)
        im, lab = self._interp_points(grid, True, self.labels is not None)
        return im, lab

