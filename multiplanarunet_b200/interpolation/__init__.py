from .sample_grid import (sample_plane_at, plane_basis, plane_basis_batch, get_voxel_axes_real_space, get_pix_dim_from_affine,
                          sample_random_views_with_angle_restriction, get_angle, view_offsets, plane_mgrid,
                          mgrid_to_points, points_to_mgrid)  # noqa: F401
from .view_interpolator import ViewInterpolator  # noqa: F401
