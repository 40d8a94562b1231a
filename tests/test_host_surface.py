"""CPU tests of the reference-facing host surface: NIfTI IO, YAMLHParams, CLI dispatcher / flags,
project-dir helpers, host geometry."""
import os

import numpy as np
import pytest

from multiplanarunet_b200.hyperparameters import YAMLHParams
from multiplanarunet_b200.image.nifti import read_nifti, write_nifti

PRESET = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multiplanarunet_b200", "bin",
                      "defaults", "MultiPlanar", "train_hparams.yaml")


def test_nifti_round_trip(tmp_path):
    rng = np.random.RandomState(0)
    aff = np.diag([1.0, 0.5, 0.1, 1.0])
    aff[:3, 3] = [3.0, -2.0, 7.5]
    for dt in (np.float32, np.uint8, np.int16):
        a = (rng.rand(12, 14, 16, 3) * 100).astype(dt)
        for ext in (".nii", ".nii.gz"):
            p = str(tmp_path / ("x" + ext))
            write_nifti(p, a, aff)
            b, aff2, hdr = read_nifti(p)
            assert b.dtype == a.dtype and np.array_equal(a, b) and np.allclose(aff, aff2)
    # geometry the reference's own integration test pins (tests/integration/test_image_pair_with_valid_image.py:86-108)
    shape = np.array([12, 14, 16])
    pix = np.linalg.norm(aff[:3, :3], axis=0)
    assert np.allclose((shape - 1) / 2, [5.5, 6.5, 7.5])
    assert np.allclose(shape * pix, [12, 7, 1.6])


def test_yaml_hparams_preserves_text_and_sets_values(tmp_path):
    p = tmp_path / "train_hparams.yaml"
    p.write_text(open(PRESET).read())
    hp = YAMLHParams(str(p))
    assert set(hp) >= {"train_data", "val_data", "test_data", "aug_data", "build", "fit"}
    assert not any(k.startswith("__CB") for k in hp)
    assert hp["build"]["complexity_factor"] == 2 and hp["fit"]["optimizer_kwargs"]["lr"] == 5e-5
    assert hp.get_from_anywhere("scaler") == "RobustScaler" and hp.get_from_anywhere("nope", 7) == 7
    assert [c["nickname"] for c in hp["fit"]["callbacks"]] == ["rlop", "tb", "mcp_clean", "es", "timer", "csv"]
    assert hp.set_value("build", "dim", 192)
    assert not hp.set_value("build", "depth", 9)          # existing non-null value is kept
    assert hp.set_value("build", "depth", 3, overwrite=True)
    with pytest.raises(AttributeError):
        hp.set_value("build", "not_a_key", 1)
    hp.save_current()
    txt = p.read_text()
    assert "# Callback presets referenced from fit.callbacks" in txt  # comments survive
    hp2 = YAMLHParams(str(p))
    assert hp2["build"]["dim"] == 192 and hp2["build"]["depth"] == 3 and hp2["build"]["n_classes"] is None


def test_cli_dispatch_and_flags(tmp_path):
    from multiplanarunet_b200.bin import mp, predict, train, train_fusion
    parser = mp.get_parser()
    for script in ("init_project", "train", "train_fusion", "predict", "predict_3D"):
        assert parser.parse_args([script, "--x"]).script == script
    a = train.get_argparser().parse_args(["--project_dir", "p", "--num_GPUs", "2", "--overwrite", "--no_val",
                                          "--train_images_per_epoch", "10", "--epochs", "3"])
    assert a.num_GPUs == 2 and a.overwrite and a.no_val and a.epochs == 3 and a.val_images_per_epoch == 3500
    a = predict.get_argparser().parse_args(["-f", "x.nii", "--sum_fusion", "--no_argmax", "--out_dir", "o", "--continue"])
    assert a.f == "x.nii" and a.sum_fusion and a.no_argmax and a.continue_
    a = train_fusion.get_argparser().parse_args([])
    assert a.batch_size == 2 ** 17 and a.epochs == 30 and a.images_per_round == 5 and a.dice_weight == "uniform"
    mp.entry_func(["init_project", "--name", "proj", "--root", str(tmp_path), "--data_dir", "/data/set"])
    hp = YAMLHParams(str(tmp_path / "proj" / "train_hparams.yaml"))
    assert hp["train_data"]["base_dir"] == "/data/set/train" and hp["aug_data"]["base_dir"] == "/data/set/aug"
    with pytest.raises(RuntimeError):
        train.validate_project_dir(str(tmp_path / "nope"))
    with pytest.raises(NotImplementedError):
        mp.entry_func(["predict_3D"])


def test_best_and_last_model_selection(tmp_path):
    from multiplanarunet_b200.utils.utils import get_best_model, get_last_model, pred_to_class
    d = tmp_path / "model"
    d.mkdir()
    with pytest.raises(OSError):
        get_best_model(str(d))
    (d / "model_weights.npz").write_bytes(b"")
    assert get_best_model(str(d)).endswith("model_weights.npz")
    for ep, v in ((3, 0.71234), (12, 0.69999), (7, 0.80001)):
        (d / ("@epoch_%02d_val_dice_%.5f.npz" % (ep, v))).write_bytes(b"")
    assert "val_dice_0.80001" in get_best_model(str(d))
    path, ep = get_last_model(str(d))
    assert ep == 12 and "@epoch_12" in path
    p = np.random.RandomState(0).rand(4, 5, 6, 3).astype(np.float32)
    assert np.array_equal(pred_to_class(p), p.argmax(-1).astype(np.uint8))


def test_unet_filter_rule_and_param_count():
    from multiplanarunet_b200.models.unet import unet_filters
    from oracle.unet import count_params, init_params
    assert unet_filters(4, 2) == [90, 181, 362, 724, 1448]   # int(64 * 2**i * sqrt(2)), unet.py:91,120
    assert unet_filters(4, 1) == [64, 128, 256, 512, 1024]
    assert count_params(init_params(5, 1, 4, 2.0)) == 62050512  # 62.05 M trainable at cf=2


def test_augmenter_registry_and_argument_errors():
    """Elastic2D is looked up by name from the YAML (sequences/utils.py:38-47) and validates its arguments like
    the reference (augmentation/augmenters.py:41-55); without a GPU the call itself fails loudly."""
    import pytest
    from multiplanarunet_b200 import augmentation
    aug = augmentation.__dict__["Elastic2D"](alpha=[0, 450], sigma=[20, 30], apply_prob=0.333)
    assert str(aug) == "Elastic2D(alpha=[0, 450], sigma=[20, 30], apply_prob=0.333)"
    assert aug.weight == 0.33
    for bad in (dict(alpha=[1, 2, 3], sigma=1, apply_prob=0.5), dict(alpha=[5, 1], sigma=1, apply_prob=0.5),
                dict(alpha=1, sigma=[3, 3], apply_prob=0.5), dict(alpha=1, sigma=2, apply_prob=1.5)):
        with pytest.raises(ValueError):
            augmentation.Elastic2D(**bad)
    from multiplanarunet_b200.augmentation.elastic_deformation import gaussian_taps
    w, r = gaussian_taps(25.0)
    assert r == 100 and len(w) == 201 and abs(w.sum() - 1) < 1e-12 and np.array_equal(w, w[::-1])
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            augmentation.elastic_transform_2d(np.zeros((8, 8), np.float32), None, 1.0, 1.0)


def test_get_sequence_registry():
    """sequences.get_sequence (mpunet/sequences/utils.py:5-79): iso_live -> 2D live-view sequence with augmenter
    objects for training only; 3D styles refused by name; unknown styles are a ValueError like the reference."""
    import pytest
    from multiplanarunet_b200.sequences import IsotrophicLiveViewSequence2D, get_sequence
    views = np.eye(3)
    kw = dict(views=views, sample_dim=32, real_space_span=30.0, n_classes=3, batch_size=4, intrp_style="iso_live")
    aug = [dict(cls_name="Elastic2D", kwargs=dict(alpha=[0, 450], sigma=[20, 30], apply_prob=0.333))]
    msgs = []
    tr = get_sequence([object()], is_validation=False, logger=msgs.append, augmenters=aug, **kw)
    va = get_sequence([object()], is_validation=True, logger=msgs.append, augmenters=aug, **kw)
    assert isinstance(tr, IsotrophicLiveViewSequence2D) and len(tr.list_of_augmenters) == 1
    assert va.list_of_augmenters is None and va.noise_sd == 0.0 and tr.noise_sd == 0.1
    assert tr.n_fg_slices == 2 and tr.force_all_fg            # batch 4 > 2 foreground classes
    assert any("augmenters" in str(m) for m in msgs)
    with pytest.raises(NotImplementedError):
        get_sequence([object()], False, **dict(kw, intrp_style="iso_live_3d"))
    with pytest.raises(ValueError):
        get_sequence([object()], False, **dict(kw, intrp_style="bogus"))
    with pytest.raises(NotImplementedError):
        get_sequence([object()], False, augmenters=[dict(cls_name="Elastic3D", kwargs={})], **kw)


def test_plane_basis_batch_is_bit_identical_to_plane_basis():
    """The vectorised basis construction used for the 320 candidate planes of a training batch equals the scalar
    mirror of sample_grid.py:192-224 bit for bit, including the |n| < 0.2 flip and the flat-normal special case."""
    import numpy as np
    from multiplanarunet_b200.interpolation import plane_basis, plane_basis_batch
    rng = np.random.RandomState(3)
    n = 3000
    views = rng.randn(n, 3)
    views[:40] = [0, 0, 1]
    views[40:80] = [0.1, 0.15, 0.9]
    views[80:120] = [-0.3, -0.1, 0.9]
    noise = rng.normal(scale=0.1, size=(n, 3))
    noise[:20] = 0
    B = plane_basis_batch(views, noise)
    for i in range(n):
        assert np.array_equal(B[i], plane_basis(views[i], noise[i].copy())), i
