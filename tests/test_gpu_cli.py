"""End-to-end CLI round trip on a toy project (GPU): init_project -> train -> train_fusion -> predict,
checking the project-dir artefacts the reference's scripts produce."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _toy(shape, seed):
    rng = np.random.RandomState(seed)
    ax = [np.arange(s, dtype=np.float32) for s in shape]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    lab = np.zeros(shape, np.uint8)
    img = rng.randn(*shape).astype(np.float32) * 5
    for c in (1, 2):
        ctr = rng.uniform(0.3, 0.7, 3) * np.array(shape)
        r = rng.uniform(0.18, 0.28) * min(shape)
        m = (X - ctr[0]) ** 2 + (Y - ctr[1]) ** 2 + (Z - ctr[2]) ** 2 < r * r
        lab[m] = c
        img[m] += 40.0 * c
    return img, lab


def test_cli_round_trip(tmp_path):
    from multiplanarunet_b200.bin import mp
    from multiplanarunet_b200.hyperparameters import YAMLHParams
    from multiplanarunet_b200.image import read_nifti, write_nifti
    data = tmp_path / "data"
    n = 0
    for split, k in (("train", 2), ("val", 1), ("test", 1)):
        for sub in ("images", "labels"):
            os.makedirs(data / split / sub)
        for i in range(k):
            img, lab = _toy((48, 48, 48), n)
            n += 1
            write_nifti(str(data / split / "images" / ("im%d.nii.gz" % i)), img, np.eye(4))
            write_nifti(str(data / split / "labels" / ("im%d.nii.gz" % i)), lab, np.eye(4))
    mp.entry_func(["init_project", "--name", "proj", "--root", str(tmp_path), "--data_dir", str(data)])
    proj = str(tmp_path / "proj")
    hp = YAMLHParams(os.path.join(proj, "train_hparams.yaml"))
    hp.set_value("build", "complexity_factor", 0.125, overwrite=True)
    hp.set_value("fit", "batch_size", 8, overwrite=True)
    hp.save_current()
    mp.entry_func(["train", "--project_dir", proj, "--overwrite", "--epochs", "2",
                   "--train_images_per_epoch", "48", "--val_images_per_epoch", "16"])
    hp = YAMLHParams(os.path.join(proj, "train_hparams.yaml"))
    assert hp["build"]["dim"] == 128 and hp["build"]["n_classes"] == 3 and hp["build"]["n_channels"] == 1
    assert os.path.exists(os.path.join(proj, "views.npz")) and np.load(os.path.join(proj, "views.npz"))["arr_0"].shape == (6, 3)
    assert os.path.exists(os.path.join(proj, "model", "model_weights.npz"))
    assert os.path.exists(os.path.join(proj, "logs", "training.csv"))
    import csv
    rows = list(csv.DictReader(open(os.path.join(proj, "logs", "training.csv"))))
    assert [r["epoch"] for r in rows] == ["0", "1"]
    assert {"loss", "lr", "val_dice", "val_precision", "val_recall", "epoch_minutes", "train_hours"} <= set(rows[0])
    ckpts = [f for f in os.listdir(os.path.join(proj, "model")) if f.startswith("@epoch_")]
    assert len(ckpts) == 1 and ckpts[0].endswith(".npz") and "_val_dice_" in ckpts[0]   # ModelCheckPointClean
    # --continue_training resumes after the last checkpoint's epoch with the logged LR (models/model_init.py:25-51)
    mp.entry_func(["train", "--project_dir", proj, "--continue_training", "--epochs", "3",
                   "--train_images_per_epoch", "48", "--val_images_per_epoch", "16"])
    rows = list(csv.DictReader(open(os.path.join(proj, "logs", "training.csv"))))
    assert [r["epoch"] for r in rows] == ["0", "1", "2"]
    mp.entry_func(["train_fusion", "--project_dir", proj, "--epochs", "2", "--batch_size", "65536"])
    fdir = os.path.join(proj, "model", "fusion_weights")
    assert len(os.listdir(fdir)) == 1
    out = str(tmp_path / "preds")
    mp.entry_func(["predict", "--project_dir", proj, "--out_dir", out])
    pred, aff, _ = read_nifti(os.path.join(out, "nii_files", "im0_PRED.nii.gz"))
    assert pred.shape == (48, 48, 48) and pred.dtype == np.uint8 and pred.max() <= 2
    assert os.path.exists(os.path.join(out, "csv", "results.csv"))
    # --continue skips what is already there; --save_input_files writes per-image folders
    mp.entry_func(["predict", "--project_dir", proj, "--out_dir", out, "--continue"])
    assert len(open(os.path.join(out, "csv", "results.csv")).read().strip().splitlines()) == 2
    out3 = str(tmp_path / "preds_inputs")
    mp.entry_func(["predict", "--project_dir", proj, "--out_dir", out3, "--sum_fusion", "--save_input_files"])
    sub = os.path.join(out3, "nii_files", "im0")
    assert sorted(os.listdir(sub)) == ["im0_IMAGE.nii.gz", "im0_LABELS.nii.gz", "im0_PRED.nii.gz"]
    out2 = str(tmp_path / "preds_sum")
    mp.entry_func(["predict", "--project_dir", proj, "--out_dir", out2, "--sum_fusion", "--no_eval"])
    pred2, _, _ = read_nifti(os.path.join(out2, "nii_files", "im0_PRED.nii.gz"))
    assert pred2.shape == (48, 48, 48)
