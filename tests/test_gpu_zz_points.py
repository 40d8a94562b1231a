"""ViewInterpolator.__call__ / intrp_image / intrp_labels on explicit grids (`mpu_interp_points`) against the
outputs of the unmodified reference interpolator (tests/golden/sampler_*.npz hold exactly `interp(grid)` for grids
from sample_plane_at) - bit-exact image and labels, including the non-axis-aligned affine case (host-side
apply_rotation) - and against the plane sampler kernel on the same planes.

Validated on hardware in round 1 (GPUTEST_r01: passed); any failure is a hard failure."""
import os

import numpy as np
import pytest

import golden_inputs as gi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _check():
    from multiplanarunet_b200.interpolation import ViewInterpolator, plane_basis, plane_mgrid
    for case in gi.SAMPLER_CASES:
        z = np.load(os.path.join(GOLD, "sampler_%s.npz" % case["name"]))
        vol, lab, affine, bg = gi.sampler_volume(case)
        vi = ViewInterpolator(vol, lab, affine, bg_value=bg, bg_class=0)
        k = 0
        for view in case["views"]:
            basis = plane_basis(view, 0.)
            ref_im, ref_lb = vi.sample_planes(basis, case["offsets"], case["dim"], case["span"])
            for j, off in enumerate(case["offsets"]):
                grid, _ = plane_mgrid(basis, case["dim"], case["span"], off)
                im, lb = vi(grid)
                assert im.shape == z["im"][k].shape and im.dtype == np.float32 and lb.dtype == np.uint8
                assert np.array_equal(lb, z["lab"][k])
                assert np.array_equal(im, z["im"][k]), float(np.abs(im - z["im"][k]).max())
                assert np.array_equal(im, ref_im[j].cpu().numpy()) and np.array_equal(lb, ref_lb[j].cpu().numpy())
                assert np.array_equal(vi.intrp_image(grid), im) and np.array_equal(vi.intrp_labels(grid), lb)
                k += 1
    # empty grid and a grid entirely outside the volume
    vi = ViewInterpolator(np.zeros((4, 4, 4, 1), np.float32), np.ones((4, 4, 4), np.uint8), np.eye(4),
                          bg_value=[-3.0], bg_class=7)
    far = np.full((3, 5, 6, 1), 1e3)
    im, lb = vi(far)
    assert im.shape == (5, 6, 1) and np.all(im == -3.0) and np.all(lb == 7)
    im, lb = vi(np.zeros((3, 0, 4, 1)))
    assert im.shape == (0, 4, 1) and lb.shape == (0, 4)


def test_interp_points_matches_reference_goldens():
    _check()
