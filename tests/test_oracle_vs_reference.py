"""Pins the oracle against the LIVE reference source where it is present (this container only; the
GPU box has no /root/reference, so these tests skip there).  Wider sweep than the golden fixtures."""
import os

import numpy as np
import pytest

from conftest import has_reference
from oracle import fusion, sampler

pytestmark = pytest.mark.skipif(not has_reference(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_shim
    return ref_shim.modules()


def test_sampler_equals_reference(ref):
    rng = np.random.RandomState(3)
    vol = rng.randn(20, 16, 24, 2).astype(np.float32)
    lab = rng.randint(0, 5, size=(20, 16, 24)).astype(np.uint8)
    aff = np.eye(4)
    ivi = ref.view_interpolator.ViewInterpolator(vol, lab, aff, bg_value=[0.5, -2.0], bg_class=0)
    for view in [(0.3, 0.5, 0.8), (0, 0, 1), (0, 1, 0), (0.1, 0.1, 0.98), (-0.4, -0.4, 0.2)]:
        basis = sampler.plane_basis(view)
        for off in [-11.0, -0.37, 6.5]:
            grid, g, inv = ref.sample_grid.sample_plane_at(view, sample_dim=28, real_space_span=26,
                                                           offset_from_center=off, noise_sd=0., test_mode=True)
            im_r, lab_r = ivi(grid)
            pts = sampler.plane_points(basis, 28, 26, off)
            assert np.array_equal(pts, np.moveaxis(grid[..., 0], 0, -1))
            im_o, lab_o = sampler.sample_plane(vol, lab, (1, 1, 1), basis, 28, 26, off, [0.5, -2.0], 0)
            assert np.array_equal(im_o, im_r) and np.array_equal(lab_o, lab_r)
            assert np.array_equal(np.linalg.inv(basis), inv)


def test_noisy_basis_equals_reference(ref):
    for seed in range(5):
        np.random.seed(seed)
        grid, g, inv = ref.sample_grid.sample_plane_at((0.3, 0.5, 0.8), 8, 10, 1.0, noise_sd=0.1, test_mode=True)
        np.random.seed(seed)
        noise = np.random.normal(scale=0.1, size=3)
        assert np.array_equal(np.linalg.inv(sampler.plane_basis((0.3, 0.5, 0.8), noise)), inv)


def test_mapping_equals_reference(ref):
    if ref.fuse_and_predict is None:
        pytest.skip("fuse_and_predict not importable under the shim")
    rng = np.random.RandomState(4)
    shape = (14, 18, 12)

    class Img:
        pass
    img = Img()
    img.shape = shape + (1,)
    img.affine = np.eye(4)
    vg_r = ref.sample_grid.get_voxel_grid_real_space(img)
    vg_o = fusion.voxel_grid_real_space(shape, np.eye(3))
    assert np.array_equal(vg_r, vg_o)
    pred = rng.rand(20, 20, 26, 3).astype(np.float32)
    g = np.linspace(-9, 9, 20)
    for view in [(0.3, 0.5, 0.8), (1, 0, 0)]:
        ib = np.linalg.inv(sampler.plane_basis(view))
        offs = sampler.view_offsets(20, 18, 26)
        a = ref.fuse_and_predict.map_real_space_pred(pred, (g, g, offs), ib, vg_r)
        b = fusion.map_real_space_pred(pred, (g, g, offs), ib, vg_o)
        assert np.array_equal(a, b)


def test_host_geometry_mirror_equals_reference(ref):
    """The product's host-side plane geometry (multiplanarunet_b200.interpolation) against the reference."""
    from multiplanarunet_b200.interpolation import sample_plane_at, view_offsets
    for view in [(0.3, 0.5, 0.8), (0, 0, 1), (-0.7, 0.1, 0.2)]:
        (basis, off), g, inv = sample_plane_at(view, 32, 30, 2.5, 0., test_mode=True)
        grid, g_r, inv_r = ref.sample_grid.sample_plane_at(view, 32, 30, 2.5, 0., test_mode=True)
        assert np.array_equal(inv, inv_r) and np.array_equal(g, g_r) and off == 2.5
        from multiplanarunet_b200.interpolation import plane_mgrid
        dense, g_d = plane_mgrid(basis, 32, 30, 2.5)          # the grid ViewInterpolator.__call__ consumes
        assert dense.shape == grid.shape == (3, 32, 32, 1) and np.array_equal(dense, grid) and np.array_equal(g_d, g_r)
    assert np.array_equal(view_offsets(64, 64.0, "same+20"), sampler.view_offsets(64, 64.0, "same+20"))


def test_dice_all_equals_reference():
    """evaluate/metrics.py imports tensorflow at module level; the shim's stub satisfies the import and
    dice/dice_all themselves are plain numpy."""
    from oracle import metrics as om, ref_shim
    ref_shim.install()
    import importlib.util
    import os
    path = os.path.join(ref_shim.REF_ROOT, "mpunet", "evaluate", "metrics.py")
    spec = importlib.util.spec_from_file_location("_ref_metrics", path)
    rm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rm)
    rng = np.random.RandomState(5)
    for n_classes, kw in [(5, {}), (5, dict(ignore_zero=False)), (None, {}), (4, dict(skip_if_no_y=True)),
                          (3, dict(smooth=0.5))]:
        k = n_classes or 6
        a = rng.randint(0, k, size=(9, 10, 11)).astype(np.uint8)
        b = rng.randint(0, k, size=(9, 10, 11)).astype(np.uint8)
        a[a == 2] = 0  # a class missing from y_true
        want = rm.dice_all(a, b, n_classes=n_classes, **kw)
        got = om.dice_all(a, b, n_classes=n_classes, **kw)
        assert np.array_equal(want, got, equal_nan=True)
    assert rm.dice(a > 1, b > 2) == om.dice(a > 1, b > 2)


def test_elastic_augmenter_equals_reference():
    """Elastic2D batch call: same mask / alpha / sigma / noise draw order as augmentation/augmenters.py:87-109."""
    from oracle import elastic, ref_shim
    ref_shim.install()
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ra = importlib.import_module("mpunet.augmentation.augmenters")
    rng = np.random.RandomState(7)
    xs = [rng.randn(24, 20, 1).astype(np.float32) for _ in range(6)]
    ys = [rng.randint(0, 3, size=(24, 20)).astype(np.uint8) for _ in range(6)]
    bgs = [[0.1 * i] for i in range(6)]
    np.random.seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rx, ry, rw = ra.Elastic2D(alpha=[0, 60], sigma=[2, 4], apply_prob=0.5)(
            batch_x=[x.copy() for x in xs], batch_y=[y.copy() for y in ys], bg_values=bgs, batch_w=[1.0] * 6)
    np.random.seed(3)
    ox, oy, ow = elastic.Elastic2D(alpha=[0, 60], sigma=[2, 4], apply_prob=0.5)(
        [x.copy() for x in xs], [y.copy() for y in ys], bgs, [1.0] * 6, rng=np.random)
    assert rw == ow and 0.33 in ow and 1.0 in ow
    for a, b in zip(rx, ox):
        assert np.array_equal(a, b)
    for a, b in zip(ry, oy):
        assert np.array_equal(a, b)


def test_unet_graph_is_the_references_own():
    """The reference's UNet class, executed unmodified under oracle/keras_shim.py, against the oracle's layer table and
    the product's filter rule over a sweep of configurations (names, shapes, parameter count, receptive field)."""
    from multiplanarunet_b200.models.unet import unet_filters
    from oracle import keras_shim
    from oracle.unet import count_params, init_params, layer_specs
    RefUNet = keras_shim.reference_unet_class()
    for n_classes, n_channels, depth, cf, dim in [(3, 1, 4, 0.125, 32), (5, 1, 4, 2.0, 64), (2, 3, 4, 1.0, 64),
                                                  (9, 2, 3, 0.5, 48), (4, 1, 2, 1.0, 16)]:
        m = RefUNet(n_classes=n_classes, dim=dim, n_channels=n_channels, depth=depth, complexity_factor=cf,
                    logger=lambda *a, **k: None)
        ref = [(l.name, {k: tuple(v.shape) for k, v in l.weights.items()}) for l in m.layers if l.weights]
        P = init_params(n_classes, n_channels, depth, cf)
        assert [r[0] for r in ref] == [s[0] for s in layer_specs(n_classes, n_channels, depth, cf)]
        for name, shapes in ref:
            assert {k: tuple(v.shape) for k, v in P[name].items()} == shapes
        assert m.trainable_count() == count_params(P)
        enc = [m.get_layer("encoder_L%d_conv1" % i).filters for i in range(depth)] + [m.get_layer("bottom_conv1").filters]
        assert enc == unet_filters(depth, cf)
        assert not m.label_crop.any() and m.img_shape == (dim, dim, n_channels)
    # the concatenation order of the up blocks is [skip, upsampled] (unet.py:167-169)
    cat = m.get_layer("upsample_L0_concat")
    assert [n.layer.name for n in cat.output.inputs] == ["encoder_L%d_BN" % (depth - 1), "upsample_L0_BN1"]


def test_fusion_model_is_the_references_own():
    """The reference's FusionModel / FusionLayer / reg / sparse_generalized_dice_loss, executed unmodified under
    oracle/keras_shim.py, against the oracle on a sweep of shapes (incl. a 2-class, 3-view case and [N] labels)."""
    from oracle import fusion, keras_shim
    lf, fm = keras_shim.reference_fusion_modules()
    for n, V, C, seed in [(512, 6, 5, 0), (300, 3, 2, 1), (64, 1, 4, 2)]:
        rng = np.random.RandomState(seed)
        x = rng.dirichlet(np.ones(C), size=(n, V)).astype(np.float32)
        y = rng.randint(0, C, size=(n, 1)).astype(np.uint8)
        W = rng.uniform(0.5, 1.5, (V, C)).astype(np.float32)
        b = (0.1 * rng.randn(1, C)).astype(np.float32)
        M = fm.FusionModel(n_inputs=V, n_classes=C, weight="uniform", logger=lambda *a, **k: None, verbose=False)
        lay = M.layers[-1]
        assert M.count_params() == V * C + C and lay.weights["W"].shape == (V, C) and lay.weights["b"].shape == (1, C)
        assert float(lay.W.min()) == 1.0 and float(np.abs(lay.b).max()) == 0.0     # constant(1) / constant(0) init
        lay.W[...] = W
        lay.b[...] = b
        p_ref = M.predict(x)
        p = fusion.fusion_forward(x, W, b)
        assert np.abs(p_ref - p).max() < 1e-6 and np.array_equal(p_ref.argmax(-1), p.argmax(-1))
        loss, _, _ = fusion.gdl_loss_and_grads(x, y, W, b, reg=1e-6)
        p64 = lay.call(keras_shim._t(x, np.float64))
        ref_loss = float(M.loss(y, p64)) + sum(lay.regularization_losses())
        assert abs(ref_loss - loss) < 1e-7
        assert abs(float(M.loss(y.reshape(-1), p64)) - float(M.loss(y, p64))) < 1e-12   # [N] and [N, 1] labels


def test_auditor_equals_reference(tmp_path):
    """The reference's Auditor (image/auditor.py:73-260, unmodified; nibabel.load replaced by an adapter over NIfTI files
    laid out from the specification) against the package's: sample dim, real-space span, channels, classes - incl. a
    case where the span is shrunk (nearest valid dim < 0.9 x span / resolution) and multi-channel volumes."""
    import sys
    import types
    from oracle import ref_shim
    from test_image_pair_cpu import nifti1_bytes
    from multiplanarunet_b200.image import Auditor
    from multiplanarunet_b200.image.nifti import read_nifti
    ref_shim.install()

    class _Hdr(dict):
        def get_zooms(self):
            return tuple(self["pixdim"][1:4])

    class _Img(object):
        def __init__(self, path):
            self._data, self.affine, h = read_nifti(path)
            self.shape = self._data.shape
            self.header = _Hdr(pixdim=np.asarray(h["pixdim"], dtype=np.float32))

        def get_data_dtype(self):
            return self._data.dtype

        def get_data(self):
            return self._data

        get_fdata = get_data

    nib = sys.modules["nibabel"]
    saved = nib.load
    nib.load = _Img
    try:
        import importlib
        ref_auditor = importlib.import_module("mpunet.image.auditor")
        rng = np.random.RandomState(5)
        cases = [
            ([((96, 110, 80), (1.0, 1.0, 1.5)), ((128, 100, 90), (0.9, 0.9, 1.2))], 1),
            ([((300, 280, 40), (0.5, 0.5, 4.0)), ((256, 256, 36), (0.6, 0.6, 4.5)), ((320, 300, 44), (0.45, 0.45, 4.0))], 1),
            ([((64, 64, 64), (2.0, 2.0, 2.0))], 3),
            ([((700, 650, 600), (1.0, 1.0, 1.0))], 1),      # span / res far above max_dim: the span is shrunk
        ]
        for ci, (vols, n_ch) in enumerate(cases):
            ims, labs = [], []
            for vi, (shape, pix) in enumerate(vols):
                small = tuple(max(4, s // 16) for s in shape)       # keep the files tiny: scale voxel size up instead
                pix_s = tuple(p * s / q for p, s, q in zip(pix, shape, small))
                data = rng.randn(*(small + ((n_ch,) if n_ch > 1 else ()))).astype(np.float64)
                aff = np.diag(list(pix_s) + [1.0])
                p_im = tmp_path / ("c%d_im%d.nii" % (ci, vi))
                p_im.write_bytes(nifti1_bytes(data, aff))
                lab = rng.randint(0, 4 + ci, size=small).astype(np.float64)
                lab.flat[:4 + ci] = np.arange(4 + ci)                # every class present
                p_lab = tmp_path / ("c%d_lab%d.nii" % (ci, vi))
                p_lab.write_bytes(nifti1_bytes(lab, aff))
                ims.append(str(p_im))
                labs.append(str(p_lab))
            ref = ref_auditor.Auditor(ims, labs, logger=lambda *a, **k: None)
            mine = Auditor(ims, labs)
            assert mine.sample_dim_2D == int(ref.sample_dim_2D), ci
            assert abs(mine.real_space_span_2D - float(ref.real_space_span_2D)) < 1e-4 * float(ref.real_space_span_2D), ci
            assert mine.n_channels == ref.n_channels == n_ch and mine.n_classes == ref.n_classes == 4 + ci
    finally:
        nib.load = saved


def test_view_sampling_and_model_selection_equal_reference(ref, tmp_path):
    """Host helpers on the path of `mp train` / `mp predict`, against the reference functions themselves:
    sample_random_views_with_angle_restriction (sample_grid.py:133-173; same RNG consumption => identical views.npz for a
    seed), get_best_model / get_last_model (utils/utils.py:88-130) over checkpoint-name layouts."""
    import importlib
    from multiplanarunet_b200.interpolation import sample_random_views_with_angle_restriction
    from multiplanarunet_b200.utils.utils import get_best_model, get_last_model
    for seed, n, ang in [(0, 6, 60), (1, 6, 60), (2, 3, 75), (3, 9, 40)]:
        np.random.seed(seed)
        want = ref.sample_grid.sample_random_views_with_angle_restriction(n, ang, logger=lambda *a, **k: None)
        np.random.seed(seed)
        got = sample_random_views_with_angle_restriction(n, ang)
        assert np.array_equal(got, want)
    ru = importlib.import_module("mpunet.utils")
    layouts = [
        ["@epoch_03_val_dice_0.71230.h5", "@epoch_10_val_dice_0.80011.h5", "@epoch_07_val_dice_0.79000.h5"],
        ["@epoch_02_val_loss_0.91000.h5", "@epoch_05_val_loss_0.35000.h5"],
        ["@epoch_04_dice_0.50000.h5", "@epoch_09_dice_0.45000.h5"],
        ["@epoch_01_loss_1.25000.h5", "@epoch_12_loss_0.75000.h5", "model_weights.h5"],
        ["model_weights.h5"],
    ]
    for i, names in enumerate(layouts):
        d = tmp_path / ("m%d" % i)
        d.mkdir()
        for nme in names:
            (d / nme).write_bytes(b"x")
        assert os.path.basename(get_best_model(str(d))) == os.path.basename(ru.get_best_model(str(d)))
        mine, theirs = get_last_model(str(d)), ru.get_last_model(str(d))
        assert os.path.basename(mine[0]) == os.path.basename(theirs[0]) and int(mine[1]) == int(theirs[1])
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(OSError):
        get_best_model(str(empty))
    with pytest.raises(OSError):
        ru.get_best_model(str(empty))


def test_validation_metrics_equal_reference():
    """compute_dice against the reference's Validation._compute_dice (callbacks/validation.py:60-90, loaded from its own
    file with a stub for tensorflow.keras.callbacks.Callback), bit for bit, incl. classes without relevant / selected
    pixels; and the confusion counts of its counting thread (:91-131) against the oracle's integer counts."""
    import queue
    import sys
    import threading
    import types
    from collections import defaultdict
    from oracle import keras_shim, metrics as ometrics
    from multiplanarunet_b200.evaluate.metrics import compute_dice
    keras_shim.install()
    cb = types.ModuleType("tensorflow.keras.callbacks")
    cb.Callback = type("Callback", (object,), {})
    sys.modules["tensorflow.keras.callbacks"] = cb
    sys.modules["tensorflow.keras"].callbacks = cb
    em = keras_shim._load_reference_file("mpunet/evaluate/metrics.py", "_ref_evaluate_metrics")
    pkg = sys.modules.setdefault("mpunet.evaluate", types.ModuleType("mpunet.evaluate"))
    pkg.dice_all = em.dice_all
    sys.modules["mpunet.evaluate.metrics"] = em
    val = keras_shim._load_reference_file("mpunet/callbacks/validation.py", "_ref_callbacks_validation")
    rng = np.random.RandomState(9)
    for n_classes in (2, 5, 9):
        rel = rng.randint(0, 5000, n_classes).astype(np.uint64)
        sel = rng.randint(0, 5000, n_classes).astype(np.uint64)
        rel[rng.randint(n_classes)] = 0
        sel[rng.randint(n_classes)] = 0
        tp = np.minimum(rel, sel) // 2
        want = val.Validation._compute_dice(tp=tp, rel=rel, sel=sel)
        got = compute_dice(tp=tp, rel=rel, sel=sel)
        for w, g in zip(want, got):
            assert g.dtype == w.dtype == np.float32 and np.array_equal(g, w)
    # the counting thread of the reference on two batches == the oracle's counts
    n_classes = 4
    q = queue.Queue()
    preds = [rng.rand(3, 16, 16, n_classes).astype(np.float32) for _ in range(2)]
    trues = [rng.randint(0, n_classes, (3, 16, 16, 1)).astype(np.uint8) for _ in range(2)]
    for p, t in zip(preds, trues):
        q.put(([p], [t]))
    TPs, relv, selv = (defaultdict(lambda: np.zeros(n_classes, np.uint64)) for _ in range(3))
    val.Validation._count_cm_elements_from_queue(q, 2, TPs, relv, selv, ["t"], [n_classes], threading.Lock())
    tp2, rel2, sel2 = ometrics.cm_counts(np.concatenate([t.ravel() for t in trues]),
                                        np.concatenate([p.argmax(-1).ravel() for p in preds]), n_classes)
    assert np.array_equal(TPs["t"], tp2) and np.array_equal(relv["t"], rel2) and np.array_equal(selv["t"], sel2)


def test_cli_flags_and_defaults_equal_reference():
    """Every `mp` script on the path exposes exactly the reference's command-line surface: the same option strings,
    argparse action types and defaults (mpunet/bin/{train,predict,train_fusion,init_project,predict_3D}.py parsers,
    imported from the reference under the shim and compared field by field)."""
    import importlib
    import sys
    import types
    from oracle import keras_shim
    keras_shim.install()
    # imported at module level by reference modules the scripts pull in, not needed by the parsers
    for top, sub, attrs in (("ruamel", "yaml", {"YAML": type("YAML", (object,), {})}),
                            ("matplotlib", "pyplot", {})):
        if top not in sys.modules:
            m_sub = types.ModuleType(top + "." + sub)
            m_sub.__dict__.update(attrs)
            m_sub.__getattr__ = lambda attr: (lambda *a, **k: None)
            m_top = types.ModuleType(top)
            m_top.__path__ = []
            m_top.use = lambda *a, **k: None
            setattr(m_top, sub, m_sub)
            sys.modules[top], sys.modules[top + "." + sub] = m_top, m_sub

    def surface(parser):
        return {(a.option_strings[0] if a.option_strings else a.dest): (tuple(a.option_strings), type(a).__name__,
                                                                        a.default, a.nargs, a.type)
                for a in parser._actions if a.dest != "help"}

    checked = 0
    for name, fn in [("train", "get_argparser"), ("predict", "get_argparser"), ("train_fusion", "get_argparser"),
                     ("init_project", "get_parser"), ("predict_3D", "get_argparser")]:
        try:
            ref_mod = importlib.import_module("mpunet.bin." + name)
        except Exception as e:  # a model family outside the path failed to import under the shim
            if name in ("train", "predict"):
                raise
            continue
        mine = importlib.import_module("multiplanarunet_b200.bin." + name)
        want = surface(getattr(ref_mod, fn)())
        got = surface(getattr(mine, fn if hasattr(mine, fn) else "get_argparser")())
        assert got == want, (name, {k: (want.get(k), got.get(k)) for k in set(want) | set(got) if want.get(k) != got.get(k)})
        checked += 1
    assert checked == 5, checked


def test_output_bias_initialisation_equals_reference():
    """set_bias_weights (utils/utils.py:205-242: bias of the softmax layer from class frequencies, biased_output_layer of
    train_hparams.yaml) against the reference function, on given class counts and on counts estimated from label maps."""
    import contextlib
    import importlib
    from multiplanarunet_b200.utils.utils import set_bias_weights
    from oracle import ref_shim
    ref_shim.install()
    if not hasattr(np, "int"):
        np.int = int  # utils.py:221 (removed in numpy 1.24; the reference pins an older numpy)
    ru = importlib.import_module("mpunet.utils.utils")

    def softmax():
        pass

    class Lay(object):
        activation = softmax

        def __init__(self, n):
            self.w = [np.zeros((1, 1, 8, n), np.float32), np.zeros(n, np.float32)]

        def get_weights(self):
            return [w.copy() for w in self.w]

        def set_weights(self, ws):
            self.w = [np.asarray(w) for w in ws]

    class Img(object):
        def __init__(self, lab):
            self.labels = lab

    class Queue(object):                         # what the reference's estimate path needs from a data queue
        def __init__(self, images):
            self.dataset, self._i = images, 0

        @contextlib.contextmanager
        def get_random_image(self):
            im = self.dataset[self._i % len(self.dataset)]
            self._i += 1
            yield im

    quiet = lambda *a, **k: None
    for counts in ([700, 120, 90, 60, 30], [1, 1], [10 ** 9, 5, 7]):
        a, b = Lay(len(counts)), Lay(len(counts))
        ru.set_bias_weights(a, None, class_counts=np.asarray(counts), logger=quiet)
        set_bias_weights(b, None, class_counts=np.asarray(counts), logger=quiet)
        assert np.allclose(a.w[-1], b.w[-1], rtol=1e-6, atol=1e-7) and abs(np.linalg.norm(b.w[-1]) - 1.0) < 1e-6
    rng = np.random.RandomState(2)
    images = [Img(rng.randint(0, 4, (6, 7, 5)).astype(np.uint8)) for _ in range(3)]
    a, b = Lay(4), Lay(4)
    ru.set_bias_weights(a, Queue(images), logger=quiet)
    set_bias_weights(b, images, logger=quiet)
    assert np.allclose(a.w[-1], b.w[-1], rtol=1e-6, atol=1e-7)
