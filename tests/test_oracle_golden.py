"""The CPU oracle reproduces the committed golden vectors, which were produced by the UNMODIFIED
reference source (oracle/make_golden.py).  Bit-exact for images (float32), labels and gathers."""
import os

import numpy as np

import golden_inputs as gi
from oracle import fusion, sampler

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_sampler_matches_reference_goldens():
    for case in gi.SAMPLER_CASES:
        z = np.load(os.path.join(GOLD, "sampler_%s.npz" % case["name"]))
        vol, lab, affine, bg = gi.sampler_volume(case)
        k = 0
        for view in case["views"]:
            basis = sampler.plane_basis(view)
            for off in case["offsets"]:
                im, lb = sampler.sample_plane(vol, lab, case["pix"], basis, case["dim"], case["span"], off, bg, 0)
                assert np.array_equal(np.linalg.inv(basis), z["inv_basis"][k])
                assert np.array_equal(im, z["im"][k]), (case["name"], view, off)
                assert np.array_equal(lb, z["lab"][k]), (case["name"], view, off)
                hd = case["span"] // 2
                assert np.array_equal(np.linspace(-hd, hd, case["dim"]), z["axis"][k])
                k += 1


def test_mapping_matches_reference_goldens():
    for case in gi.MAPPING_CASES:
        z = np.load(os.path.join(GOLD, "mapping_%s.npz" % case["name"]))
        preds, grids, inv_bases, shape, affine = gi.mapping_inputs(case)
        vg = fusion.voxel_grid_real_space(shape, affine[:3, :3])
        assert np.array_equal(vg[:, :2, :2, :2], z["vgrid_corner"])
        oob = 0.0
        for v, (p, g, ib) in enumerate(zip(preds, grids, inv_bases)):
            mapped = fusion.map_real_space_pred(np.moveaxis(p, 0, 2), g, ib, vg)
            assert np.array_equal(mapped, z["mapped"][v].astype(np.float32))
            oob += (mapped[..., 0] == 1.0).mean()
        # the oblique views leave the volume corners outside the sampled stack -> one-hot background
        assert 0.0 < oob / len(preds) < 0.9


def test_view_offsets_and_plane_axis():
    offs = sampler.view_offsets(256, 256.0, "same+20")
    assert len(offs) == 276 and offs[0] == -offs[-1]
    res = 256.0 / 255
    assert np.isclose(offs[-1], (256.0 + 20 * res) / 2)
    a = sampler.plane_axis(256, 256.0)
    assert a[0] == -128.0 and np.isclose(a[-1], 128.0)


def test_fusion_formulas():
    rng = np.random.RandomState(0)
    x = rng.rand(100, 6, 5).astype(np.float32)
    W = np.ones((6, 5), np.float32)
    b = np.zeros(5, np.float32)
    p = fusion.fusion_forward(x, W, b)
    assert np.allclose(p.sum(-1), 1, atol=1e-6)
    # with unit weights the fused argmax equals the argmax of the plain view sum (monotone softmax)
    assert np.array_equal(p.argmax(-1), x.sum(1).argmax(-1))
    # analytic gradient vs central differences (float64)
    y = rng.randint(0, 5, 100)
    W = rng.uniform(0.5, 1.5, (6, 5))
    b = 0.1 * rng.randn(5)
    loss, dW, db = fusion.gdl_loss_and_grads(x, y, W, b)
    eps = 1e-6
    for (i, j) in [(0, 0), (3, 2), (5, 4)]:
        Wp, Wm = W.copy(), W.copy()
        Wp[i, j] += eps
        Wm[i, j] -= eps
        num = (fusion.gdl_loss_and_grads(x, y, Wp, b)[0] - fusion.gdl_loss_and_grads(x, y, Wm, b)[0]) / (2 * eps)
        assert abs(num - dW[i, j]) < 1e-7
    bp, bm = b.copy(), b.copy()
    bp[1] += eps
    bm[1] -= eps
    num = (fusion.gdl_loss_and_grads(x, y, W, bp)[0] - fusion.gdl_loss_and_grads(x, y, W, bm)[0]) / (2 * eps)
    assert abs(num - db[1]) < 1e-7


def test_dice_all():
    a = np.array([0, 1, 1, 2, 2, 2])
    b = np.array([0, 1, 2, 2, 2, 0])
    d = fusion.dice_all(a, b, 3)
    assert np.allclose(d, [(1 + 2 * 1) / (1 + 2 + 1), (1 + 2 * 2) / (1 + 3 + 3)])


def test_validation_counts_and_dice():
    """oracle/metrics.py: the bincount rule of callbacks/validation.py:117-131 against brute force, and
    _compute_dice's zero-fill convention."""
    from oracle import metrics as om
    rng = np.random.RandomState(2)
    k = 4
    y = rng.randint(0, k, size=500)
    scores = rng.rand(500, k).astype(np.float32)
    scores[:50] = 0.25  # ties: first maximum wins
    p = scores.argmax(-1)
    tp, rel, sel = om.cm_counts(y, scores, k)
    for c in range(k):
        assert tp[c] == np.sum((y == c) & (p == c))
        assert rel[c] == np.sum(y == c)
        assert sel[c] == np.sum(p == c)
    assert np.all(p[:50] == 0)
    pr, rc, dc = om.compute_dice(tp, rel, sel)
    for c in range(k):
        assert abs(dc[c] - 2 * tp[c] / float(rel[c] + sel[c])) < 1e-6
    pr, rc, dc = om.compute_dice(np.array([0, 3]), np.array([0, 4]), np.array([0, 3]))
    assert pr[0] == rc[0] == dc[0] == 0 and abs(dc[1] - 2 * 1.0 * 0.75 / 1.75) < 1e-6


def test_elastic_matches_reference_goldens():
    """oracle/elastic.py against tests/golden/elastic.npz (outputs of the unmodified reference function with
    numpy's global generator seeded): image bit-exact, labels bit-exact."""
    import golden_inputs as gi
    from oracle import elastic
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "elastic.npz"))
    for case in gi.ELASTIC_CASES:
        im, lab = gi.elastic_inputs(case)
        np.random.seed(case["seed"])
        o, l = elastic.elastic_transform_2d(im, lab, case["alpha"], case["sigma"], case["bg"], rng=np.random)
        assert np.array_equal(o, z["im_" + case["name"]]), case["name"]
        assert np.array_equal(l, z["lab_" + case["name"]]), case["name"]
    # the strong case must exercise the out-of-bounds fill
    assert np.any(z["im_strong"] == np.float32(0.5))


def test_view_stack_matches_reference_goldens():
    """oracle.sampler.get_view_from + robust_scale against the reference's IsotrophicLiveViewSequence2D.get_view_from
    (7 threads, reference ViewInterpolator, MultiChannelScaler(RobustScaler)) - tests/golden/view_stack_*.npz."""
    for case in gi.VIEW_STACK_CASES:
        z = np.load(os.path.join(GOLD, "view_stack_%s.npz" % case["name"]))
        vol, lab, affine, bg = gi.view_stack_inputs(case)
        pix = np.linalg.norm(affine[:3, :3], axis=0)
        rot = None
        if case["affine"] == "rot":
            rot = np.diag(pix).dot(np.linalg.inv(affine[:3, :3]))
        X, y, grid, inv_basis = sampler.get_view_from(vol, lab, pix, case["view"], case["dim"], case["span"], bg, 0,
                                                      z["center"], z["scale"], case["n_planes"], rot_mat=rot)
        assert X.dtype == np.float32 and np.array_equal(X, z["X"]), case["name"]
        assert np.array_equal(y, z["y"])
        assert np.array_equal(grid[0], z["axis"]) and np.array_equal(grid[2], z["offsets"])
        assert np.array_equal(inv_basis, z["inv_basis"])


def test_batch_rules_match_reference_goldens():
    """The acceptance decisions of the reference's _get_valid_slice_from over a fixed candidate list
    (tests/golden/batch_rules_*.npz) against (a) the oracle restatement and (b) the PRODUCT's host-side rule loop
    (IsotrophicLiveViewSequence2D.select_slices), both fed per-candidate facts computed by the numpy oracle."""
    from multiplanarunet_b200.sequences.isotrophic_live_view_sequence_2d import IsotrophicLiveViewSequence2D
    saw_invalid_im = saw_no_fg = saw_retry = False
    for case in gi.BATCH_RULE_CASES:
        z = np.load(os.path.join(GOLD, "batch_rules_%s.npz" % case["name"]))
        vol, lab, views, cand_view, cand_off, cand_noise, bg = gi.batch_rule_inputs(case)
        B, T = cand_view.shape
        nfg = case["n_classes"] - 1
        present = np.zeros((B, T, nfg), bool)
        valid = np.zeros((B, T), bool)
        for s in range(B):
            for t in range(T):
                basis = sampler.plane_basis(views[cand_view[s, t]], cand_noise[s, t])
                im, lb = sampler.sample_plane(vol, lab, (1, 1, 1), basis, case["dim"], case["span"], cand_off[s, t], bg, 0)
                present[s, t] = np.isin(np.arange(1, case["n_classes"]), lb)
                valid[s, t] = sampler.is_valid_im(im, bg)
        saw_invalid_im |= not valid.all()
        saw_no_fg |= not present.any(-1).all()
        saw_retry |= bool((z["picks"] > 0).any())
        n_fg_slices = int(np.ceil(B * case["fg_frac"]))
        picks, counts = sampler.select_slices(present, valid, B, n_fg_slices, B > nfg, nfg)
        assert np.array_equal(picks, z["picks"]) and np.array_equal(counts, z["fg_counts"]), case["name"]
        seq = IsotrophicLiveViewSequence2D([], views=views, sample_dim=case["dim"], real_space_span=case["span"],
                                           n_classes=case["n_classes"], batch_size=B, fg_batch_fraction=case["fg_frac"])
        picks2, count2 = seq.select_slices(present, valid)
        assert np.array_equal(picks2, z["picks"]) and count2 == z["fg_counts"][-1], case["name"]
        # accepted planes: scaled image and labels equal the reference's batch
        from multiplanarunet_b200.sequences.isotrophic_live_view_sequence_2d import robust_scaler_stats
        cen, scl = robust_scaler_stats(vol)
        for s in range(B):
            t = picks[s]
            basis = sampler.plane_basis(views[cand_view[s, t]], cand_noise[s, t])
            im, lb = sampler.sample_plane(vol, lab, (1, 1, 1), basis, case["dim"], case["span"], cand_off[s, t], bg, 0,
                                          cen, scl)
            assert np.array_equal(im, z["x"][s]) and np.array_equal(lb, z["y"][s])
    assert saw_invalid_im and saw_no_fg and saw_retry  # the cases exercise both rejection reasons


def test_pairwise_tree_reproduces_numpy_sum_and_reference_centre():
    """interpolation/voxel_center.py: the recursion tree of numpy's pairwise_sum, and (with leaf sums taken on the
    host here) the exact grid centre the reference subtracts (tests/golden/voxel_center.npz)."""
    from multiplanarunet_b200.interpolation.voxel_center import combine, pairwise_tree
    z = np.load(os.path.join(GOLD, "voxel_center.npz"))

    def leaf_sum(a):
        n = len(a)
        if n < 8:
            res = 0.0
            for v in a:
                res += v
            return res
        r = list(a[:8])
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += a[i]
            i += 1
        return res

    for k, (shape, kind) in enumerate(gi.VOXEL_CENTER_CASES):
        if np.prod(shape) > 300000:
            continue  # pure-Python leaves: keep the CPU suite fast (the GPU test covers every case)
        A = gi.voxel_center_affine(kind)[:3, :3]
        grid = np.mgrid[0:shape[0]:1, 0:shape[1]:1, 0:shape[2]:1]
        pts = np.stack([g.ravel() for g in grid], 1)
        real = A.dot(pts.T).T
        n = len(pts)
        ls, ll, _, _, _ = pairwise_tree(n)
        mean = np.array([combine(n, np.array([leaf_sum(real[s:s + l, r].tolist()) for s, l in zip(ls, ll)])) / n
                         for r in range(3)])
        assert np.array_equal(mean, z["mean_%d" % k]), (shape, kind)


def test_unet_graph_matches_reference_goldens():
    """tests/golden/unet_graph_*.npz were produced by running the reference's OWN UNet.init_model (mpunet/models/
    unet.py:114-216, unmodified) under oracle/keras_shim.py: the oracle's layer table must be that graph - names in
    creation order, kernel / BN shapes, parameter counts (62 050 512 trainable at the benchmark configuration) - and its fp32
    inference output must equal the reference graph's, executed by the shim's numpy layers on the same seeded weights."""
    import json
    from oracle.unet import UNetOracle, count_params, init_params, layer_specs
    for name, kw in gi.UNET_GRAPH_CASES.items():
        z = np.load(os.path.join(GOLD, "unet_graph_%s.npz" % name))
        ref_layers = [l for l in json.loads(str(z["layers"])) if l["shapes"]]
        specs = layer_specs(kw["n_classes"], kw["n_channels"], kw["depth"], kw["complexity_factor"])
        assert [l["name"] for l in ref_layers] == [s[0] for s in specs]
        P = init_params(kw["n_classes"], kw["n_channels"], kw["depth"], kw["complexity_factor"], seed=1, randomize_bn=True)
        for l in ref_layers:
            assert {k: list(v.shape) for k, v in P[l["name"]].items()} == l["shapes"], l["name"]
        assert count_params(P) == int(z["trainable_params"])
        n_bn = sum(P[l["name"]]["gamma"].size for l in ref_layers if "gamma" in l["shapes"])
        assert int(z["count_params"]) == count_params(P) + 2 * n_bn   # Keras counts the moving statistics too
        assert not z["label_crop"].any()
        if "probs" in z.files:
            got = UNetOracle(kw["n_classes"], kw["n_channels"], kw["depth"], kw["complexity_factor"],
                             params=P).predict(gi.unet_graph_input(kw))
            assert got.shape == z["probs"].shape
            assert np.abs(got - z["probs"]).max() < 1e-6
    assert int(np.load(os.path.join(GOLD, "unet_graph_benchmark.npz"))["trainable_params"]) == 62050512


def test_fusion_layer_and_dice_loss_match_reference_goldens():
    """tests/golden/fusion_ref.npz: the reference's own FusionLayer.call, regulariser and sparse_generalized_dice_loss
    (fusion_model.py:9-43, loss_functions.py:207-246) executed unmodified under oracle/keras_shim.py.  The oracle's
    restatement must give the same probabilities / labels, the same loss, and its ANALYTIC gradients must equal the
    central differences of the reference objective."""
    from oracle import fusion
    z = np.load(os.path.join(GOLD, "fusion_ref.npz"))
    x, y, W, b = gi.fusion_inputs()
    probs = fusion.fusion_forward(x, W, b)
    assert np.abs(probs - z["probs"]).max() < 1e-6 and np.array_equal(probs.argmax(-1).astype(np.uint8), z["labels"])
    assert int(z["n_weights"]) == W.size + b.size == 35
    loss, dW, db = fusion.gdl_loss_and_grads(x, y, W, b, reg=0.0)
    for weight in ("uniform", "simple", "square"):   # rank-2 inputs: every class weight is 1 whatever the type
        assert abs(loss - float(z["loss_" + weight])) < 1e-9
    assert np.abs(dW - z["dW_uniform_fd"]).max() < 1e-8 and np.abs(db.reshape(1, -1) - z["db_uniform_fd"]).max() < 1e-8
    loss_r, _, _ = fusion.gdl_loss_and_grads(x, y, W, b, reg=1e-6)
    assert abs((loss_r - loss) - float(z["reg"])) < 1e-12
