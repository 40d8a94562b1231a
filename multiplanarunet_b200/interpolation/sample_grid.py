"""Host-side plane geometry, mirroring mpunet/interpolation/sample_grid.py.

Only the tiny per-plane quantities are computed here (a 3x3 basis and an offset per plane, exactly as
`sample_plane_at` derives them, sample_grid.py:192-224); the per-pixel grid, the interpolation and
the scaling run in the CUDA sampler (csrc/volume_ops.cu).  The reference materialises a float64
[3,dim,dim,1] grid per plane instead (sample_grid.py:227-239).
"""
from itertools import combinations

import numpy as np


def _rotation_about(axis, angle_deg):
    """Rotation matrix in the quaternion form the reference uses (interpolation/linalg.py:33-51)."""
    theta = np.deg2rad(angle_deg)
    axis = np.asarray(axis).ravel()
    axis = axis / np.linalg.norm(axis)
    qa = np.cos(theta / 2.0)
    qb, qc, qd = -axis * np.sin(theta / 2.0)
    return np.array([[qa * qa + qb * qb - qc * qc - qd * qd, 2 * (qb * qc + qa * qd), 2 * (qb * qd - qa * qc)],
                     [2 * (qb * qc - qa * qd), qa * qa + qc * qc - qb * qb - qd * qd, 2 * (qc * qd + qa * qb)],
                     [2 * (qb * qd + qa * qc), 2 * (qc * qd - qa * qb), qa * qa + qd * qd - qb * qb - qc * qc]])


def plane_basis(norm_vector, noise_sd=0.):
    """Basis [u v n_hat] (float64, columns) of the sampling plane for a view normal; `noise_sd` is the
    reference's noise_sd (scalar: drawn here with np.random.normal like sample_grid.py:199-203, or a
    ready 3-vector)."""
    n_hat = np.array(norm_vector, np.float32)
    n_hat /= np.linalg.norm(n_hat)
    if type(noise_sd) is not np.ndarray:
        noise_sd = np.random.normal(scale=noise_sd, size=3)
    n_hat += noise_sd
    n_hat /= np.linalg.norm(n_hat)
    if np.all(n_hat[:-1] < 0.2):
        n_hat[:-1] = np.abs(n_hat[:-1])
    if np.all(np.isclose(n_hat[:-1], 0)):
        u, v = np.array([1, 0, 0]), np.array([0, 1, 0])
    else:
        tilted = n_hat.copy()
        tilted[-1] = tilted[-1] + 1
        tilted /= np.linalg.norm(tilted)
        u = _rotation_about(np.cross(n_hat, tilted), -90).dot(n_hat)
        v = np.cross(n_hat, u)
    return np.column_stack((u, v, n_hat))


def _row_norms(x):
    """np.linalg.norm of every row, bit-identical to the per-vector call (both go through the BLAS dot of the numpy
    build: a batched matmul reproduces it, an elementwise sum of squares does not in ~10 % of the rows)."""
    return np.sqrt(np.matmul(x[:, None, :], x[:, :, None]).reshape(-1))


def plane_basis_batch(norm_vectors, noise):
    """plane_basis for n planes at once: norm_vectors [n, 3], noise [n, 3] float64 (the vector added to n_hat) ->
    bases [n, 3, 3] float64, bit-identical to n calls of plane_basis (tests/test_host_surface.py).  320 candidate
    planes of a training batch take ~0.3 ms instead of ~55 ms."""
    nv = np.asarray(norm_vectors)
    n = nv.shape[0]
    n_hat = np.array(nv, np.float32)
    n_hat /= _row_norms(n_hat)[:, None]
    n_hat += np.asarray(noise, dtype=np.float64)
    n_hat /= _row_norms(n_hat)[:, None]
    flip = np.all(n_hat[:, :-1] < 0.2, axis=1)
    n_hat[flip, :-1] = np.abs(n_hat[flip, :-1])
    flat = np.all(np.isclose(n_hat[:, :-1], 0), axis=1)
    tilted = n_hat.copy()
    tilted[:, -1] = tilted[:, -1] + 1
    tilted /= _row_norms(tilted)[:, None]
    axis = np.cross(n_hat, tilted)
    with np.errstate(invalid="ignore", divide="ignore"):
        axis = axis / _row_norms(axis)[:, None]
    theta = np.deg2rad(-90)
    qa = np.cos(theta / 2.0)
    q = -axis * np.sin(theta / 2.0)
    qb, qc, qd = q[:, 0], q[:, 1], q[:, 2]
    R = np.empty((n, 3, 3), dtype=np.float64)
    R[:, 0, 0] = qa * qa + qb * qb - qc * qc - qd * qd
    R[:, 0, 1] = 2 * (qb * qc + qa * qd)
    R[:, 0, 2] = 2 * (qb * qd - qa * qc)
    R[:, 1, 0] = 2 * (qb * qc - qa * qd)
    R[:, 1, 1] = qa * qa + qc * qc - qb * qb - qd * qd
    R[:, 1, 2] = 2 * (qc * qd + qa * qb)
    R[:, 2, 0] = 2 * (qb * qd + qa * qc)
    R[:, 2, 1] = 2 * (qc * qd - qa * qb)
    R[:, 2, 2] = qa * qa + qd * qd - qb * qb - qc * qc
    u = np.matmul(R, n_hat.astype(np.float64)[:, :, None]).reshape(n, 3)
    v = np.cross(n_hat, u)
    u[flat] = (1, 0, 0)
    v[flat] = (0, 1, 0)
    out = np.empty((n, 3, 3), dtype=np.float64)
    out[:, :, 0] = u
    out[:, :, 1] = v
    out[:, :, 2] = n_hat
    return out


def sample_plane_at(norm_vector, sample_dim, real_space_span, offset_from_center, noise_sd,
                    test_mode=False):
    """Mirror of sample_grid.py:192-244 that returns the plane PARAMETERS instead of the dense grid:
    (basis, offset) and, in test_mode, also the in-plane axis and inv(basis) like the reference."""
    basis = plane_basis(norm_vector, noise_sd)
    if test_mode:
        hd = real_space_span // 2
        return (basis, float(offset_from_center)), np.linspace(-hd, hd, sample_dim), np.linalg.inv(basis)
    return basis, float(offset_from_center)


def plane_mgrid(basis, sample_dim, real_space_span, offset_from_center):
    """The dense float64 grid [3, dim, dim, 1] the reference's sample_plane_at returns (sample_grid.py:227-239):
    `basis . (a, b, offset)` over np.mgrid[-hd:hd:dim*1j] with hd = span // 2.  Only needed to feed
    ViewInterpolator.__call__ / intrp_* (the sampler kernel builds these coordinates itself)."""
    hd = real_space_span // 2
    g = np.linspace(-hd, hd, sample_dim)
    j = complex(sample_dim)
    grid = np.mgrid[-hd:hd:j, -hd:hd:j, offset_from_center:offset_from_center:1j]
    points = mgrid_to_points(grid)
    real = basis.dot(points.T).T
    return points_to_mgrid(real, grid.shape[1:]), g


def mgrid_to_points(mgrid):
    """[N, D1, D2, D3] mesh grid (or tuple) -> [D1*D2*D3, N] points (interpolation/linalg.py:5-14)."""
    pts = np.empty(shape=(int(np.prod(mgrid[0].shape)), len(mgrid)), dtype=mgrid[0].dtype)
    for k in range(len(mgrid)):
        pts[:, k] = mgrid[k].ravel()
    return pts


def points_to_mgrid(points, grid_shape):
    """Inverse of mgrid_to_points (interpolation/linalg.py:17-24)."""
    mgrid = np.empty(shape=(points.shape[1],) + tuple(grid_shape), dtype=points.dtype)
    for k in range(points.shape[1]):
        mgrid[k] = points[:, k].reshape(grid_shape)
    return mgrid


def view_offsets(sample_dim, real_space_span, n_planes="same+20", bounding_radius=None):
    """Plane offsets of an inference stack (sequences/isotrophic_live_view_sequence_2d.py:47-62)."""
    sample_res = real_space_span / (sample_dim - 1)
    if n_planes == "by_radius":
        n = int(2 * bounding_radius / sample_res)
        bounds = bounding_radius
    else:
        extra = 0
        if n_planes == "same":
            n = sample_dim
        elif isinstance(n_planes, str) and n_planes[:5] == "same+":
            extra = int(n_planes.split("+")[-1])
            n = sample_dim + extra
        else:
            n = int(n_planes)
        bounds = (real_space_span + (extra * sample_res)) / 2
    return np.linspace(-bounds, bounds, n)


def get_pix_dim_from_affine(affine):
    return np.linalg.norm(np.asarray(affine)[:3, :3], axis=0)


def get_voxel_axes_real_space(shape3, affine):
    """Centred float32 voxel axes times pixdim and the optional alignment rotation
    (sample_grid.py:63-98).  Returns ((gx, gy, gz), rot_mat | None)."""
    basis = np.asarray(affine)[:3, :3]
    pixdims = np.linalg.norm(basis, axis=0)
    transform = np.diag(pixdims)
    rot_mat = transform.dot(np.linalg.inv(basis)) if np.any(~np.isclose(transform, basis)) else None
    axes = []
    for n, pd in zip(shape3, pixdims):
        g = np.arange(n, dtype=np.float32) - np.float32((n - 1) / 2)
        axes.append((g * np.float32(pd)).astype(np.float32))
    return tuple(axes), rot_mat


def get_angle(v1, v2):
    v1_u = v1 / np.linalg.norm(v1)
    v2_u = v2 / np.linalg.norm(v2)
    return np.rad2deg(np.arccos(np.clip(np.dot(v1_u, v2_u), -1.0, 1.0)))


def sample_random_views_with_angle_restriction(views, min_angle_deg, weights=None):
    """N random unit vectors (z >= 0) whose pairwise angles exceed min_angle_deg, relaxing the bound by
    one degree per failed draw (sample_grid.py:133-173)."""
    N = views
    while True:
        dev = np.random.normal(size=(N, 3))
        v = dev / np.linalg.norm(dev, axis=1)[:, np.newaxis]
        v[:, -1] = np.abs(v[:, -1])
        if weights is not None:
            vw = v * weights
            v = vw / np.linalg.norm(vw, axis=1)[:, np.newaxis]
        angles = [get_angle(a, b) for a, b in combinations(v, 2)]
        if np.all(np.asarray(angles) > min_angle_deg):
            return v
        min_angle_deg -= 1
