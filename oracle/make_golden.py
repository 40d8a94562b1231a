"""Generate tests/golden/*.npz from the UNMODIFIED reference source (imported under oracle/ref_shim.py).

Run in the build container (the only place /root/reference exists):
    python -m oracle.make_golden
The fixtures pin the oracle (and through it the CUDA kernels) to the reference's own sampler,
interpolator and mapping code: mpunet/interpolation/{sample_grid,view_interpolator,
regular_grid_interpolator}.py and mpunet/utils/fusion/fuse_and_predict.py:92-137.
Inputs are regenerated from seeds by the tests (see tests/golden_inputs.py); only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
import golden_inputs as gi  # noqa: E402


def main():
    m = ref_shim.modules()
    sg, vi = m.sample_grid, m.view_interpolator
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # ---- sampler: planes through a small 2-channel volume, several views/offsets, with OOB regions
    for case in gi.SAMPLER_CASES:
        vol, lab, affine, bg = gi.sampler_volume(case)
        interp = vi.ViewInterpolator(vol, lab, affine, bg_value=bg, bg_class=0)
        ims, labs, inv_bases, axes = [], [], [], []
        for view in case["views"]:
            for off in case["offsets"]:
                grid, g, inv = sg.sample_plane_at(view, sample_dim=case["dim"], real_space_span=case["span"],
                                                  offset_from_center=off, noise_sd=0., test_mode=True)
                im, lb = interp(grid)
                ims.append(im)
                labs.append(lb)
                inv_bases.append(inv)
                axes.append(g)
        np.savez_compressed(os.path.join(out_dir, "sampler_%s.npz" % case["name"]),
                            im=np.stack(ims), lab=np.stack(labs), inv_basis=np.stack(inv_bases),
                            axis=np.stack(axes))
        print("sampler_%s: %d planes" % (case["name"], len(ims)))

    # ---- mapping: nearest gather of per-view prediction stacks onto the voxel grid
    for case in gi.MAPPING_CASES:
        preds, grids, inv_bases, shape, affine = gi.mapping_inputs(case)

        class Img:
            pass
        img = Img()
        img.shape = tuple(shape) + (1,)
        img.affine = affine
        vgrid = sg.get_voxel_grid_real_space(img)
        mapped = [m.fuse_and_predict.map_real_space_pred(np.moveaxis(p, 0, 2), g, ib, vgrid)
                  for p, g, ib in zip(preds, grids, inv_bases)]
        np.savez_compressed(os.path.join(out_dir, "mapping_%s.npz" % case["name"]),
                            mapped=np.stack(mapped).astype(np.float16),  # values are copies of float16-exact inputs
                            vgrid_corner=vgrid[:, :2, :2, :2])
        print("mapping_%s: %s" % (case["name"], np.stack(mapped).shape))

    # ---- Elastic2D: the reference function with numpy's global RNG seeded per case
    import importlib
    ed = importlib.import_module("mpunet.augmentation.elastic_deformation")
    outs = {}
    for case in gi.ELASTIC_CASES:
        im, lab = gi.elastic_inputs(case)
        np.random.seed(case["seed"])
        o, l = ed.elastic_transform_2d(im.copy(), lab.copy(), case["alpha"], case["sigma"], case["bg"])
        outs["im_" + case["name"]] = o
        outs["lab_" + case["name"]] = l
        print("elastic_%s: mean |delta| %.4f, %d labels moved" % (case["name"], np.abs(o - im).mean(),
                                                                 int((l != lab).sum())))
    np.savez_compressed(os.path.join(out_dir, "elastic.npz"), **outs)


if __name__ == "__main__":
    main()
