"""GPU parity tests (pytest -m gpu): every CUDA kernel family against its reference, through the C-ABI.

  * tcgen05 multi-tap GEMMs (forward / dgrad / upsample-conv / wgrad) vs torch fp32 conv2d on the same
    bf16-rounded operands (a floating-point kernel: tolerance 0.02 + 1% written below);
  * plane sampler, multi-view map+fuse, fusion training vs the numpy oracle (bit-exact gathers/labels)
    and vs the golden vectors produced by the unmodified reference (tests/golden/).
"""
import os

import numpy as np
import pytest

import golden_inputs as gi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bg():
    import bringup_gemm
    return bringup_gemm


@pytest.mark.parametrize("case", ["fwd_small_64", "fwd_small_64_relu_bias", "fwd_c96_n96", "fwd_c192_n192_mask",
                                  "fwd_two_src", "fwd_big_n256", "fwd_cin8", "upconv", "upconv_odd",
                                  "wgrad_small", "wgrad_split", "wgrad_c96", "wgrad_n256"])
def test_gemm_case(case):
    fn, kw = _bg().CASES[case]
    assert fn(**kw), case


def test_volume_kernels_match_oracle():
    import bringup_volume as bv
    assert bv.sampler_cases()
    assert bv.mapfuse_case()
    assert bv.fusion_train_case()


def test_sampler_matches_reference_goldens():
    """CUDA sampler vs outputs of the unmodified reference (bit-exact image float32 and labels)."""
    from multiplanarunet_b200.interpolation import ViewInterpolator, plane_basis
    for case in gi.SAMPLER_CASES:
        z = np.load(os.path.join(GOLD, "sampler_%s.npz" % case["name"]))
        vol, lab, affine, bg = gi.sampler_volume(case)
        vi = ViewInterpolator(vol, lab, affine, bg_value=bg, bg_class=0)
        k = 0
        for view in case["views"]:
            basis = plane_basis(view, 0.)
            im, lb = vi.sample_planes(basis, case["offsets"], case["dim"], case["span"])
            im, lb = im.cpu().numpy(), lb.cpu().numpy()
            for j in range(len(case["offsets"])):
                assert np.array_equal(lb[j], z["lab"][k])
                assert np.array_equal(im[j], z["im"][k]), float(np.abs(im[j] - z["im"][k]).max())
                k += 1


def test_mapping_matches_reference_goldens():
    import torch
    from multiplanarunet_b200.utils.fusion.fuse_and_predict import _map_fuse
    for case in gi.MAPPING_CASES:
        z = np.load(os.path.join(GOLD, "mapping_%s.npz" % case["name"]))
        preds, grids, inv_bases, shape, affine = gi.mapping_inputs(case)
        C = preds[0].shape[-1]
        _, _, combined = _map_fuse([torch.as_tensor(p).cuda() for p in preds], grids, inv_bases, shape,
                                   affine[:3, :3], np.ones((len(preds), C), np.float32), np.zeros(C, np.float32),
                                   want_labels=False, want_combined=True)
        assert np.array_equal(combined.cpu().numpy(), z["mapped"].astype(np.float32))


def test_map_fuse_linspace_path_equals_table_path():
    """mpu_map_fuse_linspace (axis values recomputed in registers, 16-byte gathers) against mpu_map_fuse (axis
    tables) on the same inputs: labels, probabilities and the per-view mapped volumes must be IDENTICAL, for every
    class count the fast path is instantiated for, with a sheared / rotated affine, odd sizes, and views whose
    stacks leave parts of the volume out of bounds."""
    import torch
    from multiplanarunet_b200.interpolation import plane_basis, view_offsets
    from multiplanarunet_b200.utils.fusion.fuse_and_predict import _map_fuse, linspace_params
    rng = np.random.RandomState(5)
    th = 0.4
    R = np.array([[np.cos(th), -np.sin(th), 0.1], [np.sin(th), np.cos(th), 0], [0, 0.05, 1]])
    for C in (2, 3, 4, 5, 7, 8):
        dim, span, n, V = 37, 33.0, 45, 3
        shape = (29, 31, 23)
        affine = R.dot(np.diag([1.0, 0.7, 1.3]))
        g = np.linspace(-(span // 2), span // 2, dim)
        views = [(0.3, 0.5, 0.8), (0, 0, 1), (-0.7, 0.1, 0.2)]
        preds = [torch.rand(n, dim, dim, C, device="cuda") for _ in views]
        grids = [(g, g, view_offsets(dim, span, n))] * V
        assert linspace_params(g) is not None and linspace_params(grids[0][2]) is not None
        ibs = [np.linalg.inv(plane_basis(v, 0.)) for v in views]
        W = rng.uniform(0.5, 1.5, (V, C)).astype(np.float32)
        b = (0.1 * rng.randn(C)).astype(np.float32)
        a = _map_fuse(preds, grids, ibs, shape, affine, W, b, want_probs=True, want_combined=True)
        t = _map_fuse(preds, grids, ibs, shape, affine, W, b, want_probs=True, want_combined=True, force_tables=True)
        for x, y in zip(a, t):
            assert torch.equal(x, y), C
        a = _map_fuse(preds, grids, ibs, shape, affine, sum_fusion=True)
        t = _map_fuse(preds, grids, ibs, shape, affine, sum_fusion=True, force_tables=True)
        assert torch.equal(a[0], t[0])
        assert 0.02 < float((a[0] == 0).float().mean()) < 0.98 or C == 2
    # systematic exact ties: integer-spaced axes, voxel centres exactly half way between nodes (tie -> lower index)
    dim, span, n, C = 33, 32.0, 33, 5
    g = np.linspace(-(span // 2), span // 2, dim)
    offs = view_offsets(dim, span, n)
    assert np.array_equal(g, np.arange(-16.0, 17.0)) and np.array_equal(offs, g)
    preds = [torch.rand(n, dim, dim, C, device="cuda") for _ in range(2)]
    ibs = [np.linalg.inv(plane_basis(v, 0.)) for v in [(0, 0, 1), (1, 0, 0)]]
    W = np.ones((2, C), np.float32)
    a = _map_fuse(preds, [(g, g, offs)] * 2, ibs, (30, 30, 30), np.eye(3), W, np.zeros(C, np.float32),
                  want_probs=True, want_combined=True)
    t = _map_fuse(preds, [(g, g, offs)] * 2, ibs, (30, 30, 30), np.eye(3), W, np.zeros(C, np.float32),
                  want_probs=True, want_combined=True, force_tables=True)
    for x, y in zip(a, t):
        assert torch.equal(x, y)
    # voxel 0 sits at -14.5: the tie goes to node -15 = index 1 on every axis of the axis-aligned view
    assert torch.equal(a[2][0][0, 0, 0], preds[0][1, 1, 1])


def test_map_fuse_full_size_properties():
    """256^3 x 6 views x 5 classes (BASELINE config 3 size): size-independent properties.
    (a) one-hot per-view predictions of a constant class fuse to that class inside the sampled region and
    to background outside; (b) the label volume is invariant to a positive common scale of W up to fp32
    near-ties; (c) with unit weights, fused labels equal sum-fusion labels (softmax is monotone)."""
    import torch
    from multiplanarunet_b200.interpolation import plane_basis, view_offsets
    from multiplanarunet_b200.utils.fusion.fuse_and_predict import _map_fuse
    dim, span, C, V = 256, 256.0, 5, 6
    np.random.seed(0)
    from multiplanarunet_b200.interpolation import sample_random_views_with_angle_restriction
    views = sample_random_views_with_angle_restriction(V, 60)
    offs = view_offsets(dim, span, "same+20")
    n = len(offs)
    g = np.linspace(-(span // 2), span // 2, dim)
    rng = torch.Generator(device="cuda").manual_seed(0)
    pred = torch.rand(n, dim, dim, C, device="cuda", generator=rng)
    pred = pred / pred.sum(-1, keepdim=True)
    preds = [pred] * V  # same buffer for every view keeps this at 1.4 GB
    grids = [(g, g, offs)] * V
    ibs = [np.linalg.inv(plane_basis(v, 0.)) for v in views]
    W = np.ones((V, C), np.float32)
    b = np.zeros(C, np.float32)
    lab1, _, _ = _map_fuse(preds, grids, ibs, (dim, dim, dim), np.eye(3), W, b)
    lab2, _, _ = _map_fuse(preds, grids, ibs, (dim, dim, dim), np.eye(3), 3.0 * W, b)
    lab3, _, _ = _map_fuse(preds, grids, ibs, (dim, dim, dim), np.eye(3), sum_fusion=True)
    assert lab1.shape == (dim, dim, dim) and lab1.dtype == torch.uint8
    # fp32 softmax can merge near-ties (first index wins), so allow a vanishing fraction of flips
    assert float((lab1 != lab2).float().mean()) < 1e-4 and float((lab1 != lab3).float().mean()) < 1e-4
    const = torch.zeros(n, dim, dim, C, device="cuda")
    const[..., 3] = 1.0
    lab4, _, comb = _map_fuse([const] * V, grids, ibs, (dim, dim, dim), np.eye(3), W, b, want_combined=False)
    # the volume centre is inside every view's stack; the far corner is outside all of them
    assert int(lab4[128, 128, 128]) == 3 and int(lab4[0, 0, 0]) == 0
    assert set(torch.unique(lab4).tolist()) <= {0, 3}


def test_label_counts_and_dice_all_exact():
    """mpu_label_counts (integer work: bit-exact) against oracle/metrics.py, labels and score inputs, an
    unaligned odd length, accumulation over calls, and dice_all on a 256^3-sized volume via the invariant
    sum(relevant) == sum(selected) == n."""
    import torch
    from multiplanarunet_b200.evaluate import compute_dice, dice, dice_all, label_counts
    from oracle import metrics as om
    rng = np.random.RandomState(11)
    k = 5
    n = 100003
    y = rng.randint(0, k, size=n).astype(np.uint8)
    p = rng.randint(0, k, size=n).astype(np.uint8)
    c = label_counts(y, p, k).cpu().numpy()
    tp, rel, sel = om.cm_counts(y, p, k)
    assert np.array_equal(c[0], tp.astype(np.int64)) and np.array_equal(c[1], rel.astype(np.int64))
    assert np.array_equal(c[2], sel.astype(np.int64))
    scores = rng.rand(n, k).astype(np.float32)
    scores[:1000] = 0.5
    acc = label_counts(y, scores, k)
    acc = label_counts(y, scores, k, counts=acc)  # accumulates
    tp, rel, sel = om.cm_counts(y, scores, k)
    assert np.array_equal(acc.cpu().numpy(), 2 * np.stack([tp, rel, sel]).astype(np.int64))
    for kw in [dict(n_classes=k), dict(n_classes=k, ignore_zero=False), dict(n_classes=None),
               dict(n_classes=7, skip_if_no_y=True)]:
        assert np.array_equal(dice_all(y, p, **kw), om.dice_all(y, p, **kw), equal_nan=True)
    assert dice(y > 1, p > 2) == om.dice(y > 1, p > 2)
    pr, rc, dc = compute_dice(tp, rel, sel)
    pr2, rc2, dc2 = om.compute_dice(tp, rel, sel)
    assert np.array_equal(dc, dc2) and np.array_equal(pr, pr2) and np.array_equal(rc, rc2)
    # full-size volume (BASELINE config 3): device-resident labels
    big_t = torch.randint(0, k, (256 ** 3,), dtype=torch.uint8, device="cuda")
    big_p = torch.randint(0, k, (256 ** 3,), dtype=torch.uint8, device="cuda")
    cb = label_counts(big_t, big_p, k)
    assert int(cb[1].sum()) == 256 ** 3 and int(cb[2].sum()) == 256 ** 3
    assert torch.equal(cb[0], torch.stack([((big_t == i) & (big_p == i)).sum() for i in range(k)]))


def test_elastic_2d_against_reference_goldens_and_oracle():
    """mpu_elastic_2d: the float64 Gaussian filter / displaced bilinear + nearest resampling against the
    reference's own outputs (tests/golden/elastic.npz) and the oracle on a 256x256 batch.
    Bar: labels bit-exact, image within 1 float32 ulp-scale (2e-6 abs on O(1) data) - the device evaluates the
    same float64 expression order, so equality is expected; the bound only allows for libm/FMA differences in
    scipy's build."""
    import os
    import torch
    import golden_inputs as gi
    from multiplanarunet_b200.augmentation import Elastic2D, elastic_transform_2d
    from oracle import elastic as oe
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "elastic.npz"))
    for case in gi.ELASTIC_CASES:
        im, lab = gi.elastic_inputs(case)
        np.random.seed(case["seed"])
        o, l = elastic_transform_2d(im, lab, case["alpha"], case["sigma"], case["bg"])
        assert o.dtype == np.float32 and l.dtype == np.uint8
        assert np.abs(o - z["im_" + case["name"]]).max() <= 2e-6, case["name"]
        assert np.array_equal(l, z["lab_" + case["name"]]), case["name"]
        print("elastic %s: image bit-exact=%s" % (case["name"], np.array_equal(o, z["im_" + case["name"]])))
    # batch call at the benchmark slice size, YAML default parameter ranges
    rng = np.random.RandomState(4)
    B, H, W = 6, 256, 256
    xs = rng.randn(B, H, W, 1).astype(np.float32)
    ys = rng.randint(0, 5, size=(B, H, W)).astype(np.uint8)
    bgs = [[-0.5]] * B
    np.random.seed(9)
    ox, oy, ow = oe.Elastic2D([0, 450], [20, 30], 0.5)(list(xs), list(ys), bgs, [1.0] * B, rng=np.random)
    np.random.seed(9)
    gx, gy, gw = Elastic2D([0, 450], [20, 30], 0.5)(torch.as_tensor(xs).cuda(), torch.as_tensor(ys).cuda(), bgs,
                                                   [1.0] * B)
    assert gw == ow and 0.33 in gw
    assert np.abs(gx.cpu().numpy() - np.stack(ox)).max() <= 2e-6
    assert np.array_equal(gy.cpu().numpy(), np.stack(oy))


def test_fusion_persistent_epoch_equals_per_batch_launches(monkeypatch):
    """mpu_fusion_train_epoch as ONE cooperative launch (grid barrier per batch, every block applies the same Adam
    update to its own copy of the 35 parameters) against the one-launch-per-batch path and against the float64 oracle
    (oracle/fusion.py: GDL gradients + Keras Adam), batch by batch, including a ragged last batch."""
    import torch
    from multiplanarunet_b200.models import FusionModel
    from oracle import fusion
    rng = np.random.RandomState(12)
    N, V, C, B = 70000, 6, 5, 8192   # 9 batches, the last one with 4464 points
    X = rng.rand(N, V, C).astype(np.float32)
    X /= X.sum(-1, keepdims=True)
    y = rng.randint(0, C, size=N).astype(np.uint8)
    W0 = rng.uniform(0.5, 1.5, size=(V, C)).astype(np.float32)
    b0 = (0.1 * rng.randn(C)).astype(np.float32)
    Xd, yd = torch.as_tensor(X).cuda(), torch.as_tensor(y).cuda()
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MPU_FUSION_EPOCH_PERSISTENT", mode)
        fm = FusionModel(V, C)
        fm.set_weights([W0, b0])
        hist = fm.fit(Xd, yd, batch_size=B, epochs=2, shuffle=False)
        torch.cuda.synchronize()
        Wm, bm = fm.get_weights()
        res[mode] = (np.asarray(Wm).copy(), np.asarray(bm).ravel().copy(), hist, fm.iterations)
    monkeypatch.delenv("MPU_FUSION_EPOCH_PERSISTENT")
    assert res["1"][3] == res["0"][3] == 18
    assert np.abs(res["1"][0] - res["0"][0]).max() < 2e-6 and np.abs(res["1"][1] - res["0"][1]).max() < 2e-6
    assert np.allclose(res["1"][2], res["0"][2], rtol=1e-6, atol=1e-9)
    # oracle: 18 sequential Adam steps in float64
    th = np.concatenate([W0.ravel(), b0]).astype(np.float64)
    m, v = np.zeros_like(th), np.zeros_like(th)
    t = 0
    for ep in range(2):
        for s in range(0, N, B):
            Wc, bc = th[:V * C].reshape(V, C), th[V * C:]
            _, dW, db = fusion.gdl_loss_and_grads(X[s:s + B], y[s:s + B], Wc.astype(np.float32), bc.astype(np.float32))
            t += 1
            th, m, v = fusion.adam_step(th, np.concatenate([dW.ravel(), db]), m, v, t, 1e-3)
    err = max(np.abs(res["1"][0].ravel() - th[:V * C]).max(), np.abs(res["1"][1] - th[V * C:]).max())
    assert err < 5e-5, err   # 18 steps of size 1e-3
    # rows in order without an index (perm == NULL): same epoch as the arange permutation
    import ctypes
    from multiplanarunet_b200 import _C
    fa, fb = FusionModel(V, C), FusionModel(V, C)
    for f in (fa, fb):
        f.set_weights([W0, b0])
    fa.fit(Xd, yd, batch_size=B, epochs=1, shuffle=False)
    fb.fit(Xd[:B], yd[:B], batch_size=B, epochs=1, shuffle=False)   # allocates the scratch buffers, one step
    fb.set_weights([W0, b0])
    fb._m.zero_()
    fb._v.zero_()
    losses = torch.zeros((N + B - 1) // B, dtype=torch.float64, device="cuda")
    _C.check(_C.lib.mpu_fusion_train_epoch(_C.ptr(Xd), _C.ptr(yd), _C.ptr(None), ctypes.c_longlong(N),
                                           ctypes.c_longlong(B), V, C, _C.ptr(fb.W), _C.ptr(fb.b), _C.ptr(fb._m),
                                           _C.ptr(fb._v), _C.ptr(fb._accum), _C.ptr(fb._counter), _C.ptr(losses),
                                           ctypes.c_float(fb.reg), ctypes.c_float(fb.lr), ctypes.c_float(fb.beta_1),
                                           ctypes.c_float(fb.beta_2), ctypes.c_float(fb.epsilon), 1,
                                           _C.current_stream()), "mpu_fusion_train_epoch")
    torch.cuda.synchronize()
    assert float((fa.W - fb.W).abs().max()) < 2e-6 and float((fa.b - fb.b).abs().max()) < 2e-6
