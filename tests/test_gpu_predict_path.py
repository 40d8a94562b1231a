"""Path-level parity of `mp predict` for one image (mpunet/bin/predict.py:294-366): sampler -> U-Net -> mapping ->
fusion -> argmax on the device against the CPU pipeline oracle (oracle/predict_pipeline.py, built from the pieces
that are pinned to the unmodified reference), on a volume with a rotated + sheared + anisotropic affine (so the
exact grid centre, apply_rotation and non-trivial voxel axes are all on the path).

  * plane stacks of every view: bit-exact (float32);
  * with the device's own per-view probabilities fed to the oracle's mapping + fusion: per-view mapped volumes
    bit-exact, fused LABEL MAP BIT-EXACT, fused probabilities within 1e-6;
  * whole pipeline against the oracle with the bf16-emulating U-Net restatement: probabilities within 2e-3 (tiny
    random-weight net, see test_gpu_variants), labels equal wherever the oracle's decision margin exceeds 4e-3.
"""
import numpy as np
import pytest

import golden_inputs as gi

pytestmark = pytest.mark.gpu


def test_predict_multi_view_label_map_matches_pipeline_oracle():
    import torch
    from multiplanarunet_b200.interpolation import sample_random_views_with_angle_restriction
    from multiplanarunet_b200.models import UNet
    from multiplanarunet_b200.sequences import IsotrophicLiveViewSequence2D, SyntheticImage
    from multiplanarunet_b200.utils.fusion.fuse_and_predict import predict_multi_view, predict_stack_device
    from oracle import predict_pipeline as pp
    from oracle.unet import UNetOracle, init_params
    rng = np.random.RandomState(7)
    shape, dim, span, K, V = (40, 36, 44), 32, 40, 3, 3
    low = rng.randn(5, 5, 6)
    vol = np.kron(low, np.ones((8, 8, 8)))[:shape[0], :shape[1], :shape[2]] + 0.2 * rng.randn(*shape)
    vol = (vol * 50 + 100).astype(np.float32)[..., None]
    affine = gi.rotated_affine()
    image = SyntheticImage(vol, None, affine, bg_value="1pct")
    np.random.seed(3)
    views = sample_random_views_with_angle_restriction(V, 60)
    P = init_params(K, 1, 4, 0.125, seed=4, randomize_bn=True)
    model = UNet(n_classes=K, dim=dim, n_channels=1, complexity_factor=0.125, max_batch=8, training=False)
    model.set_keras_weights(P)
    seq = IsotrophicLiveViewSequence2D([image], views=views, sample_dim=dim, real_space_span=span, n_classes=K,
                                       is_validation=True)
    W = rng.uniform(0.5, 1.5, (V, K)).astype(np.float32)
    b = (0.1 * rng.randn(K)).astype(np.float32)
    labels, probs, combined = predict_multi_view(model, seq, image, views, W, b, want_probs=True, want_combined=True)
    labels, probs, combined = labels.cpu().numpy(), probs.cpu().numpy(), combined.cpu().numpy()
    assert labels.shape == shape and labels.dtype == np.uint8

    # (1) device probabilities -> oracle mapping + fusion: everything downstream of the network is exact
    per_view = [predict_stack_device(model, seq, image, v, "same+20")[0].cpu().numpy() for v in views]
    lab_o, probs_o, comb_o, stacks = pp.predict_multi_view(vol, affine, views, dim, span, image.bg_value,
                                                           image.scaler_center, image.scaler_scale, W=W, b=b,
                                                           per_view_probs=per_view)
    assert np.array_equal(combined, comb_o)
    assert np.array_equal(labels, lab_o)
    assert np.abs(probs - probs_o).max() < 1e-6
    frac_oob = float((comb_o[..., 0] == 1.0).mean())
    assert 0.01 < frac_oob < 0.9 and len(np.unique(labels)) >= 2  # a non-trivial map with out-of-stack corners

    # (2) plane stacks
    for v, view in enumerate(views):
        X, _, grid, inv_basis = seq.get_view_from(image, view, "same+20")
        assert np.array_equal(X, stacks[v]), v

    # (3) whole pipeline with the restated U-Net
    oracle = UNetOracle(K, 1, 4, 0.125, params=P)
    lab_f, probs_f, _, _ = pp.predict_multi_view(vol, affine, views, dim, span, image.bg_value, image.scaler_center,
                                                 image.scaler_scale, W=W, b=b,
                                                 unet_predict=lambda x: oracle.predict(x, emulate_bf16=True))
    assert np.abs(probs - probs_f).max() < 2e-3
    top2 = np.sort(probs_f, axis=-1)[..., -2:]
    decided = (top2[..., 1] - top2[..., 0]) > 4e-3
    assert np.array_equal(labels[decided], lab_f[decided]) and decided.mean() > 0.9
    print("predict path: label map bit-exact given the device's view probabilities; vs the all-oracle pipeline "
          "max prob diff %.3g, labels differ at %.3g of voxels" % (np.abs(probs - probs_f).max(),
                                                                   float((labels != lab_f).mean())))

    # sum fusion (bin/predict.py --sum_fusion)
    lab_s, _, _ = predict_multi_view(model, seq, image, views, sum_fusion=True)
    lab_so = fusion_sum_labels(comb_o)
    assert np.array_equal(lab_s.cpu().numpy(), lab_so)


def fusion_sum_labels(combined):
    from oracle import fusion
    return fusion.merge_views(combined, sum_fusion=True)[1]
