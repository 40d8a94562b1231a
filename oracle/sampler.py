"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement in numpy of the reference's oblique-plane sampler:
  * plane basis + grid ............ mpunet/interpolation/sample_grid.py:192-244 (sample_plane_at)
  * rotation helper ............... mpunet/interpolation/linalg.py:33-51 (get_rotation_matrix)
  * voxel axes .................... mpunet/interpolation/sample_grid.py:63-98
  * index search / linear / nearest mpunet/interpolation/regular_grid_interpolator.py:204-223,252-270
  * per-channel interpolation ..... mpunet/interpolation/view_interpolator.py:62-147
  * inference plane stack ......... mpunet/sequences/isotrophic_live_view_sequence_2d.py:29-117
  * per-channel affine scaler ..... mpunet/preprocessing/scaling.py:75-88 (sklearn RobustScaler.transform)

Pinned against the unmodified reference source (imported under oracle/ref_shim.py) by
tests/test_oracle_vs_reference.py and by the committed fixtures in tests/golden/ (made by
oracle/make_golden.py).  dtype notes follow the reference exactly: plane coordinates are float64,
volume axes are float32 (so cell widths are float32 differences), interpolation weights are float64,
image output is cast to float32.
"""
import itertools

import numpy as np


def rotation_matrix(axis, angle_deg):
    """linalg.py:33-51 (quaternion form of Rodrigues)."""
    theta = np.deg2rad(angle_deg)
    axis = np.asarray(axis).ravel()
    axis = axis / np.linalg.norm(axis)
    a = np.cos(theta / 2.0)
    b, c, d = -axis * np.sin(theta / 2.0)
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    bc, ad, ac, ab, bd, cd = b * c, a * d, a * c, a * b, b * d, c * d
    return np.array([[aa + bb - cc - dd, 2 * (bc + ad), 2 * (bd - ac)],
                     [2 * (bc - ad), aa + cc - bb - dd, 2 * (cd + ab)],
                     [2 * (bd + ac), 2 * (cd - ab), aa + dd - bb - cc]])


def plane_basis(view, noise=None):
    """sample_grid.py:194-224.  `noise` = the 3-vector added to n_hat (None = zeros, i.e. noise_sd=0).
    Returns basis [u v n_hat] (float64 3x3, columns)."""
    n_hat = np.array(view, np.float32)
    n_hat /= np.linalg.norm(n_hat)
    n_hat += (np.zeros(3) if noise is None else np.asarray(noise, dtype=np.float64))
    n_hat /= np.linalg.norm(n_hat)
    if np.all(n_hat[:-1] < 0.2):
        n_hat[:-1] = np.abs(n_hat[:-1])
    if np.all(np.isclose(n_hat[:-1], 0)):
        u = np.array([1, 0, 0])
        v = np.array([0, 1, 0])
    else:
        nhat_vs = n_hat.copy()
        nhat_vs[-1] = nhat_vs[-1] + 1
        nhat_vs /= np.linalg.norm(nhat_vs)
        u = rotation_matrix(np.cross(n_hat, nhat_vs), -90).dot(n_hat)
        v = np.cross(n_hat, u)
    return np.column_stack((u, v, n_hat))


def plane_axis(dim, span):
    """In-plane coordinate axis as np.mgrid[-hd:hd:dim*1j] builds it (sample_grid.py:227-233):
    idx*step + start in float64 with hd = span // 2."""
    hd = span // 2
    start, stop = float(-hd), float(hd)
    step = (stop - start) / float(dim - 1)
    return np.arange(dim, dtype=np.float64) * step + start


def plane_points(basis, dim, span, offset):
    """Real-space coordinates of one plane: [dim, dim, 3] float64 (sample_grid.py:236-239)."""
    a = plane_axis(dim, span)
    pts = np.empty((dim * dim, 3), dtype=np.float64)
    pts[:, 0] = np.repeat(a, dim)
    pts[:, 1] = np.tile(a, dim)
    pts[:, 2] = float(offset) * 1.0
    real = basis.dot(pts.T).T
    return real.reshape(dim, dim, 3)


def voxel_axes(shape3, pixdims=(1.0, 1.0, 1.0)):
    """sample_grid.py:93-98 + :83-85.  float32 centred axes times pixdim (float32 product - the
    reference's pinned numpy 1.x keeps float32 when an array meets a float64 scalar)."""
    out = []
    for n, pd in zip(shape3, pixdims):
        g = np.arange(n, dtype=np.float32) - np.float32((n - 1) / 2)
        out.append((g * np.float32(pd)).astype(np.float32))
    return out


def find_indices(grid, x):
    """regular_grid_interpolator.py:252-270 for one axis.  Returns (i, t, oob)."""
    i = np.searchsorted(grid, x) - 1
    i[i < 0] = 0
    i[i > grid.size - 2] = grid.size - 2
    t = (x - grid[i]) / (grid[i + 1] - grid[i])
    oob = (x < grid[0]) | (x > grid[-1])
    return i, t, oob


def interp_linear(axes, values, pts, fill):
    """Trilinear gather (regular_grid_interpolator.py:204-217 + :199-200).  values [X,Y,Z] float32,
    pts [...,3] float64 -> float64 array [...]."""
    shp = pts.shape[:-1]
    p = pts.reshape(-1, 3)
    idx, ts = [], []
    oob = np.zeros(p.shape[0], dtype=bool)
    for k in range(3):
        i, t, o = find_indices(axes[k], p[:, k])
        idx.append(i)
        ts.append(t)
        oob |= o
    res = 0.0
    for edge in itertools.product(*[[i, i + 1] for i in idx]):
        w = 1.0
        for e, i, t in zip(edge, idx, ts):
            w = w * np.where(e == i, 1 - t, t)
        res = res + np.asarray(values[edge]) * w
    res[oob] = np.float32(fill)
    return res.reshape(shp)


def interp_nearest(axes, values, pts, fill):
    """Nearest gather (regular_grid_interpolator.py:219-223 + :199-200).  values [X,Y,Z(,C)]."""
    shp = pts.shape[:-1]
    p = pts.reshape(-1, 3)
    sel = []
    oob = np.zeros(p.shape[0], dtype=bool)
    for k in range(3):
        i, t, o = find_indices(axes[k], p[:, k])
        sel.append(np.where(t <= .5, i, i + 1))
        oob |= o
    res = values[tuple(sel)].copy()
    res[oob] = fill
    return res.reshape(shp + values.shape[3:])


def robust_scale(x, center, scale):
    """sklearn RobustScaler.transform on float32 data: in-place `X -= center_; X /= scale_` with
    float64 statistics -> each step computed in float64 and rounded to float32."""
    t = (x.astype(np.float64) - np.float64(center)).astype(np.float32)
    return (t.astype(np.float64) / np.float64(scale)).astype(np.float32)


def sample_plane(image, labels, pixdims, basis, dim, span, offset, bg_value, bg_class=0,
                 center=None, scale=None, rot_mat=None):
    """One plane: image [X,Y,Z,C] float32, labels [X,Y,Z] uint8 or None.
    Returns (im [dim,dim,C] float32 (scaled when center/scale given), lab [dim,dim] uint8|None)."""
    axes = voxel_axes(image.shape[:3], pixdims)
    pts = plane_points(basis, dim, span, offset)
    if rot_mat is not None:  # view_interpolator.py:54-60
        pts = rot_mat.dot(pts.reshape(-1, 3).T).T.reshape(dim, dim, 3)
    C = image.shape[-1]
    im = np.zeros((dim, dim, C), dtype=np.float32)
    for c in range(C):
        im[..., c] = interp_linear(axes, image[..., c], pts, bg_value[c])
    lab = None
    if labels is not None:
        lab = interp_nearest(axes, labels, pts, np.uint8(bg_class)).astype(np.uint8)
    if center is not None:
        for c in range(C):
            im[..., c] = robust_scale(im[..., c], center[c], scale[c])
    return im, lab


def is_valid_im(im, bg_value):
    """isotrophic_live_view_sequence.py:91-96: a slice is rejected when every channel is (np.isclose) background."""
    return any(bool(np.any(~np.isclose(im[..., i], bg))) for i, bg in enumerate(bg_value))


def select_slices(present, valid_im, batch_size, n_fg_slices, force_all_fg, n_fg_classes):
    """The reference's batch loop (isotrophic_live_view_sequence_2d.py:119-161,163-190 with the rules of
    isotrophic_live_view_sequence.py:98-128) on per-candidate facts.  present [B,T,n_fg] bool, valid_im [B,T] bool.
    Note `has_fg_vec`: __getitem__ passes its own zeros vector to every slot and _get_valid_slice_from only rebinds
    a local name, so class coverage is tracked within the tries of one slot only.  Returns (picks, fg counts)."""
    B, T = present.shape[:2]
    picks, counts = [], []
    has_fg_count = 0
    for cur_bs in range(B):
        has_fg_vec = np.zeros(n_fg_classes, dtype=np.int64)
        tries = 0
        while tries < T:
            tries += 1
            lab_present = present[cur_bs, tries - 1]
            if force_all_fg and tries < T:
                new_mask = has_fg_vec + lab_present
                if np.all(new_mask) or np.sum(new_mask == 0) < (batch_size - cur_bs):
                    has_fg_vec = new_mask
                else:
                    continue
            if np.any(lab_present):
                valid_lab, fg_change = True, 1
            elif (n_fg_slices - has_fg_count) < (batch_size - cur_bs):
                valid_lab, fg_change = True, 0
            else:
                valid_lab, fg_change = False, 0
            if valid_lab or tries == T:
                if tries == T or valid_im[cur_bs, tries - 1]:
                    has_fg_count += fg_change
                    break
        picks.append(tries - 1)
        counts.append(has_fg_count)
    return np.asarray(picks), np.asarray(counts)


def view_offsets(dim, span, n_planes="same+20"):
    """isotrophic_live_view_sequence_2d.py:47-62."""
    sample_res = span / (dim - 1)
    extra = 0
    if n_planes == "same":
        n = dim
    elif isinstance(n_planes, str) and n_planes[:5] == "same+":
        extra = int(n_planes.split("+")[-1])
        n = dim + extra
    else:
        n = int(n_planes)
    bounds = (span + (extra * sample_res)) / 2
    return np.linspace(-bounds, bounds, n)


def get_view_from(image, labels, pixdims, view, dim, span, bg_value, bg_class=0, center=None,
                  scale=None, n_planes="same+20", rot_mat=None):
    """isotrophic_live_view_sequence_2d.py:29-101.  Returns X [dim,dim,n,C] f32, y [dim,dim,n] u8|None,
    (axis, axis, offsets), inv_basis."""
    basis = plane_basis(view)
    offsets = view_offsets(dim, span, n_planes)
    n = len(offsets)
    X = np.empty((dim, dim, n, image.shape[-1]), dtype=np.float32)
    y = np.empty((dim, dim, n), dtype=np.uint8) if labels is not None else None
    for k, off in enumerate(offsets):
        im, lab = sample_plane(image, labels, pixdims, basis, dim, span, off, bg_value, bg_class,
                               center, scale, rot_mat)
        X[:, :, k, :] = im
        if y is not None:
            y[:, :, k] = lab
    hd = span // 2
    g = np.linspace(-hd, hd, dim)
    return X, y, (g, g, offsets), np.linalg.inv(basis)
