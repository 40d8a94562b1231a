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


UNET_GRAPH_CASES = [
    # name, kwargs of the reference's UNet(...), whether to store a forward pass
    ("small", dict(n_classes=3, dim=32, n_channels=1, depth=4, complexity_factor=0.125), True),
    ("rgb", dict(n_classes=2, dim=48, n_channels=3, depth=4, complexity_factor=0.25), True),
    ("benchmark", dict(n_classes=5, dim=256, n_channels=1, depth=4, complexity_factor=2.0), False),
]


def unet_graph_goldens(out_dir):
    """tests/golden/unet_graph_*.npz: the graph the reference's OWN `UNet.init_model` builds (mpunet/models/unet.py
    executed unmodified under oracle/keras_shim.py) - layer names in creation order, parameter shapes, count_params,
    receptive field, label crop - and, for the small cases, its inference output on seeded weights / inputs computed by
    the shim's numpy layers."""
    import json
    from oracle import keras_shim
    from oracle.unet import init_params
    UNet = keras_shim.reference_unet_class()
    for name, kw, forward in UNET_GRAPH_CASES:
        model = UNet(logger=lambda *a, **k: None, **kw)
        layers = [dict(name=l.name, cls=l.__class__.__name__,
                       shapes={k: list(v.shape) for k, v in l.weights.items()},
                       out=list(l.output.shape[1:])) for l in model.layers]
        rec = dict(layers=json.dumps(layers), count_params=model.count_params(),
                   trainable_params=model.trainable_count(),
                   receptive_field=np.asarray(model.receptive_field), label_crop=np.asarray(model.label_crop))
        if forward:
            P = init_params(kw["n_classes"], kw["n_channels"], kw["depth"], kw["complexity_factor"], seed=1,
                            randomize_bn=True)
            for l in model.layers:
                for k in l.weights:
                    l.weights[k] = P[l.name][k]
            x = gi.unet_graph_input(kw)
            rec["probs"] = model.predict(x)
        np.savez_compressed(os.path.join(out_dir, "unet_graph_%s.npz" % name), **rec)
        print("unet_graph_%s: %d layers, %d parameters (%d trainable)" % (name, len(layers), rec["count_params"],
                                                                         rec["trainable_params"]))


def fusion_goldens(out_dir):
    """tests/golden/fusion_ref.npz: the reference's OWN FusionModel / FusionLayer.call (mpunet/models/fusion_model.py:
    9-75), weight regulariser and sparse_generalized_dice_loss (mpunet/evaluate/loss_functions.py:207-246) executed
    unmodified under oracle/keras_shim.py (eager numpy versions of the elementary TensorFlow ops they are written in):
    probabilities, labels, the loss for the three `type_weight`s, and central-difference gradients of the reference
    objective (float64) with respect to W and b."""
    from oracle import keras_shim
    lf, fm = keras_shim.reference_fusion_modules()
    x, y, W, b = gi.fusion_inputs()
    rec = {}
    model = fm.FusionModel(n_inputs=W.shape[0], n_classes=W.shape[1], weight="uniform", logger=lambda *a, **k: None,
                           verbose=False)
    lay = model.layers[-1]
    lay.W[...] = W
    lay.b[...] = b
    probs = model.predict(x)
    rec["probs"] = probs
    rec["labels"] = probs.argmax(-1).astype(np.uint8)
    rec["n_weights"] = model.count_params()
    rec["reg"] = np.asarray(sum(lay.regularization_losses()), dtype=np.float64)

    def objective(Wv, bv, weight):
        """mean per-point loss of the reference's loss function on the reference layer's float64 output"""
        lay.weights["W"] = lay.W = keras_shim._t(Wv, np.float64)
        lay.weights["b"] = lay.b = keras_shim._t(bv, np.float64)
        p = lay.call(keras_shim._t(x, np.float64))
        per = lf.sparse_generalized_dice_loss(keras_shim._t(y), p, weight)
        return float(np.mean(per))

    W64, b64 = W.astype(np.float64), b.astype(np.float64)
    for weight in ("uniform", "simple", "square"):
        rec["loss_" + weight] = np.asarray(objective(W64, b64, weight))
    h = 1e-6
    dW, db = np.zeros_like(W64), np.zeros_like(b64)
    for idx in np.ndindex(*W64.shape):
        d = np.zeros_like(W64)
        d[idx] = h
        dW[idx] = (objective(W64 + d, b64, "uniform") - objective(W64 - d, b64, "uniform")) / (2 * h)
    for idx in np.ndindex(*b64.shape):
        d = np.zeros_like(b64)
        d[idx] = h
        db[idx] = (objective(W64, b64 + d, "uniform") - objective(W64, b64 - d, "uniform")) / (2 * h)
    rec["dW_uniform_fd"], rec["db_uniform_fd"] = dW, db
    np.savez_compressed(os.path.join(out_dir, "fusion_ref.npz"), **rec)
    print("fusion_ref: loss uniform %.6f simple %.6f square %.6f, |dW| %.3e" % (
        rec["loss_uniform"], rec["loss_simple"], rec["loss_square"], np.abs(dW).max()))


def main():
    if sys.argv[1:] == ["unet"]:  # only the network-graph fixtures
        return unet_graph_goldens(os.path.join(ROOT, "tests", "golden"))
    if sys.argv[1:] == ["fusion"]:
        return fusion_goldens(os.path.join(ROOT, "tests", "golden"))
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
    class Img:
        pass

    for case in gi.MAPPING_CASES:
        preds, grids, inv_bases, shape, affine = gi.mapping_inputs(case)
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

    # ---- exact grid centre: what get_voxel_grid_real_space subtracts (voxel 0 sits at A.0 = 0, so centred[0] = -mean)
    centers = {}
    for k, (shape, kind) in enumerate(gi.VOXEL_CENTER_CASES):
        img = Img()
        img.shape = tuple(shape) + (1,)
        img.affine = gi.voxel_center_affine(kind)
        vg = sg.get_voxel_grid_real_space(img)
        centers["mean_%d" % k] = -vg[:, 0, 0, 0]
        centers["corner_%d" % k] = vg[:, -1, -1, -1]
    np.savez_compressed(os.path.join(out_dir, "voxel_center.npz"), **centers)
    print("voxel_center: %d cases" % len(gi.VOXEL_CENTER_CASES))

    # ---- inference view stacks: the reference's IsotrophicLiveViewSequence2D.get_view_from (7 worker threads) on a
    #      duck-typed image with the reference's ViewInterpolator and MultiChannelScaler(RobustScaler)
    import importlib
    seq2d = importlib.import_module("mpunet.sequences.isotrophic_live_view_sequence_2d")
    scaling = importlib.import_module("mpunet.preprocessing.scaling")
    for case in gi.VIEW_STACK_CASES:
        vol, lab, affine, bg = gi.view_stack_inputs(case)
        image = Img()
        image.image, image.labels, image.affine, image.shape = vol, lab, affine, vol.shape
        image.n_channels, image.predict_mode = vol.shape[-1], False
        image.interpolator = vi.ViewInterpolator(vol, lab, affine, bg_value=bg, bg_class=0)
        image.scaler = scaling.get_scaler("RobustScaler").fit(vol)
        seq = object.__new__(seq2d.IsotrophicLiveViewSequence2D)
        seq.sample_dim, seq.real_space_span, seq.logger = case["dim"], case["span"], (lambda *a, **k: None)
        X, y, grid, inv_basis = seq.get_view_from(image, case["view"], case["n_planes"])
        np.savez_compressed(os.path.join(out_dir, "view_stack_%s.npz" % case["name"]), X=X, y=y, axis=grid[0],
                            offsets=grid[2], inv_basis=inv_basis,
                            center=np.array([float(s_.center_[0]) for s_ in image.scaler.scalers]),
                            scale=np.array([float(s_.scale_[0]) for s_ in image.scaler.scalers]))
        print("view_stack_%s: X %s %s, scaler centre %s" % (case["name"], X.shape, X.dtype,
                                                            [float(s_.center_[0]) for s_ in image.scaler.scalers]))

    # ---- training-batch rejection rules: the reference's _get_valid_slice_from / validate_lab(_vec) / is_valid_im
    #      with np.random replaced by a replay of the candidate list
    for case in gi.BATCH_RULE_CASES:
        vol, lab, views, cand_view, cand_off, cand_noise, bg = gi.batch_rule_inputs(case)
        image = Img()
        image.interpolator = vi.ViewInterpolator(vol, lab, np.eye(4), bg_value=bg, bg_class=0)
        image.scaler = scaling.get_scaler("RobustScaler").fit(vol)
        seq = object.__new__(seq2d.IsotrophicLiveViewSequence2D)
        seq.sample_dim, seq.real_space_span, seq.noise_sd, seq.views = case["dim"], case["span"], 0.1, views
        seq._batch_size, seq.n_classes, seq.fg_batch_fraction = case["B"], case["n_classes"], case["fg_frac"]
        seq.force_all_fg_switch, seq.fg_classes, seq.is_validation = "auto", np.arange(1, case["n_classes"]), False
        state = {"slot": 0, "try": -1}

        class Replay(object):
            @staticmethod
            def randint(lo, hi, size=None):
                state["try"] += 1
                return np.array([cand_view[state["slot"], state["try"]]])

            @staticmethod
            def uniform(lo, hi, size=None):
                return np.array([cand_off[state["slot"], state["try"]]])

            @staticmethod
            def normal(scale=1.0, size=None):
                return cand_noise[state["slot"], state["try"]].copy()

        saved = (np.random.randint, np.random.uniform, np.random.normal)
        np.random.randint, np.random.uniform, np.random.normal = Replay.randint, Replay.uniform, Replay.normal
        try:
            picks, fg_counts, ims, labs = [], [], [], []
            has_fg_count, has_fg_vec = 0, np.zeros_like(seq.fg_classes)
            for slot in range(case["B"]):
                state["slot"], state["try"] = slot, -1
                im, lb, has_fg_count = seq._get_valid_slice_from(image=image, max_tries=case["tries"],
                                                                 has_fg_vec=has_fg_vec, has_fg_count=has_fg_count,
                                                                 cur_bs=slot)
                picks.append(state["try"])
                fg_counts.append(has_fg_count)
                ims.append(im)
                labs.append(lb)
        finally:
            np.random.randint, np.random.uniform, np.random.normal = saved
        x = np.asarray(seq.scale(ims, [image.scaler] * len(ims)))
        np.savez_compressed(os.path.join(out_dir, "batch_rules_%s.npz" % case["name"]), picks=np.array(picks),
                            fg_counts=np.array(fg_counts), x=x, y=np.asarray(labs))
        print("batch_rules_%s: picks %s fg counts %s" % (case["name"], picks, fg_counts))

    # ---- Elastic2D: the reference function with numpy's global RNG seeded per case
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
    unet_graph_goldens(out_dir)
    fusion_goldens(out_dir)


if __name__ == "__main__":
    main()
