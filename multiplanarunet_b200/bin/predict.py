"""`mp predict` on the B200 engine (mirror of mpunet/bin/predict.py:19-503): per image, the six view
stacks are sampled, segmented and fused without leaving HBM (utils.fusion.predict_multi_view)."""
import csv
import os
from argparse import ArgumentParser

import numpy as np


def get_argparser():
    p = ArgumentParser(description="Predict using a mpunet model.")
    p.add_argument("--project_dir", type=str, default="./")
    p.add_argument("-f", help="Predict on a single file")
    p.add_argument("-l", help="Optional single label file to use with -f")
    p.add_argument("--dataset", type=str, default="test")
    p.add_argument("--out_dir", type=str, default="predictions")
    p.add_argument("--num_GPUs", type=int, default=1)
    p.add_argument("--sum_fusion", action="store_true")
    p.add_argument("--overwrite", action="store_true")
    p.add_argument("--no_eval", action="store_true")
    p.add_argument("--eval_prob", type=float, default=1.0)
    p.add_argument("--force_GPU", type=str, default="")
    p.add_argument("--save_input_files", action="store_true")
    p.add_argument("--no_argmax", action="store_true")
    p.add_argument("--on_val", action="store_true")
    p.add_argument("--wait_for", type=str, default="")
    p.add_argument("--continue", action="store_true", dest="continue_")
    return p


def validate_folders(base_dir, out_dir, overwrite, _continue):
    for p in ("train_hparams.yaml", "views.npz", "model"):
        if not os.path.exists(os.path.join(base_dir, p)):
            raise RuntimeError("[*] Invalid mpunet project folder: '%s'\n    Needed file/folder '%s' not found." % (base_dir, p))
    if not overwrite and not _continue and os.path.exists(out_dir) and os.listdir(out_dir):
        raise RuntimeError("[*] Output directory already exists at: %s\n  Use --overwrite to overwrite or "
                           "--continue to continue" % out_dir)


def entry_func(args=None):
    a = get_argparser().parse_args(args)
    base_dir = os.path.abspath(a.project_dir)
    out_dir = os.path.abspath(a.out_dir)
    validate_folders(base_dir, out_dir, a.overwrite, a.continue_)
    if a.force_GPU:
        os.environ["CUDA_VISIBLE_DEVICES"] = a.force_GPU
    import torch
    from .. import distributed as D
    from .. import models
    from ..hyperparameters import YAMLHParams
    from ..image import ImagePair, ImagePairLoader, write_nifti
    from ..sequences import IsotrophicLiveViewSequence2D
    from ..utils.fusion import predict_multi_view
    from ..utils.utils import get_best_model

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    D.init_from_env(device=torch.device("cuda", local))
    hp = YAMLHParams(os.path.join(base_dir, "train_hparams.yaml"))
    build, fit = dict(hp["build"]), dict(hp["fit"])
    bg_value = hp.get_from_anywhere("bg_value") or 0.0
    if a.f:
        images = [ImagePair(os.path.abspath(a.f), os.path.abspath(a.l) if a.l and not a.no_eval else None,
                            bg_value=bg_value)]
    else:
        ds = ("val" if a.on_val else a.dataset).replace("_data", "") + "_data"
        loader = ImagePairLoader(bg_value=bg_value, predict_mode=a.no_eval, **hp[ds])
        images = [loader.get(i) for i in D.shard(list(range(len(loader))))]  # independent volumes shard
    views = np.load(os.path.join(base_dir, "views.npz"))["arr_0"]
    weights = get_best_model(os.path.join(base_dir, "model"))
    # 92 planes per inference call (a 256-plane-wide stack has 276 planes = 3 x 92): fuller waves of the persistent GEMM
    # kernels on the coarse levels than with 32 (bench.py --workload predict --predict-batch)
    model = models.__dict__[build["model_class_name"]](max_batch=int(os.environ.get("MPU_PREDICT_BATCH", "92")),
                                                       training=False, **build)
    model.load_weights(weights)
    W = b = None
    if not a.sum_fusion:
        stem = os.path.splitext(os.path.basename(weights))[0]
        fbase = os.path.join(base_dir, "model", "fusion_weights", "%s_fusion_weights" % stem)
        if os.path.exists(fbase + ".npz"):
            with np.load(fbase + ".npz") as z:
                W, b = z["W"], z["b"]
        elif os.path.exists(fbase + ".h5"):  # fusion weights trained with the reference (bin/predict.py:227-233)
            from ..utils.keras_h5 import load_keras_weights
            layers = [d for d in load_keras_weights(fbase + ".h5").values() if "W" in d and "b" in d]
            if len(layers) != 1:
                raise OSError("%s.h5 does not hold exactly one fusion layer" % fbase)
            W, b = layers[0]["W"], layers[0]["b"]
        else:
            raise OSError("Fusion weights not found at %s.{npz,h5}; run `mp train_fusion` or pass --sum_fusion"
                          % fbase)
    seq = IsotrophicLiveViewSequence2D(images, views=views, sample_dim=build["dim"],
                                       real_space_span=fit["real_space_span"], n_classes=build["n_classes"],
                                       is_validation=True)
    nii_dir = os.path.join(out_dir, "nii_files")
    os.makedirs(nii_dir, exist_ok=True)
    os.makedirs(os.path.join(out_dir, "csv"), exist_ok=True)
    if a.continue_:
        # remove_already_predicted (bin/predict.py:120-131)
        done = set(f.replace("_PRED", "").split(".")[0] for f in os.listdir(nii_dir))
        if done:
            print("[OBS] Not predicting on images: {} (--continue mode)".format(sorted(done)))
        images = [im for im in images if im.identifier not in done]
    rows = []
    for image in images:
        labels, probs, _ = predict_multi_view(model, seq, image, views, W, b, sum_fusion=a.sum_fusion,
                                              want_probs=a.no_argmax)
        out = probs.cpu().numpy() if a.no_argmax else labels.cpu().numpy()
        # save_nii_files (bin/predict.py:88-117): with --save_input_files everything goes to a per-image folder
        dst = os.path.join(nii_dir, image.identifier) if a.save_input_files else nii_dir
        os.makedirs(dst, exist_ok=True)
        write_nifti(os.path.join(dst, image.identifier + "_PRED.nii.gz"), out, image.affine)
        if a.save_input_files:
            write_nifti(os.path.join(dst, image.identifier + "_IMAGE.nii.gz"), np.asarray(image.image), image.affine)
            if image.labels is not None:
                write_nifti(os.path.join(dst, image.identifier + "_LABELS.nii.gz"), np.asarray(image.labels),
                            image.affine)
        if image.labels is not None and not a.no_eval:
            if np.random.rand() > a.eval_prob:
                print("Skipping evaluation for %s... (eval_prob=%.3f)" % (image.identifier, a.eval_prob))
                continue
            from ..evaluate import dice_all
            dices = list(dice_all(image.labels, labels, n_classes=build["n_classes"], ignore_zero=True))
            rows.append([image.identifier] + [float(d) for d in dices])
            print("%s  combined dices %s  mean dice %.4f" % (image.identifier, np.round(dices, 4),
                                                             np.nanmean(dices)))
    if D.world_size() > 1:  # volumes were sharded: collect every rank's rows on rank 0
        gathered = [None] * D.world_size()
        torch.distributed.all_gather_object(gathered, rows)
        rows = [r for part in gathered for r in part]
    if rows and D.rank() == 0:
        path = os.path.join(out_dir, "csv", "results.csv")
        header = ["id"] + ["class_%d" % c for c in range(1, build["n_classes"])]
        old = []
        if a.continue_ and os.path.exists(path):
            with open(path, newline="") as f:
                old = [r for r in csv.reader(f)][1:]
        with open(path, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(header)
            w.writerows(old + rows)
