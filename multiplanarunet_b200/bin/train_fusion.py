"""`mp train_fusion` on the B200 engine (mirror of mpunet/bin/train_fusion.py:45-366): for rounds of
<= images_per_round images build the point sets X [N_vox, V, C] (per-view U-Net prediction mapped to
the voxel grid, utils/fusion/fusion_training.py:40-89) on the device and fit the 35-parameter fusion
layer with the generalized dice loss; points shard across ranks, gradients all-reduce."""
import os
from argparse import ArgumentParser

import numpy as np


def get_argparser():
    p = ArgumentParser(description="Fit a fusion model for a mpunet project")
    p.add_argument("--project_dir", type=str, default="./")
    p.add_argument("--overwrite", action="store_true")
    p.add_argument("--num_GPUs", type=int, default=1)
    p.add_argument("--images_per_round", type=int, default=5)
    p.add_argument("--batch_size", type=int, default=2 ** 17)
    p.add_argument("--epochs", type=int, default=30)
    p.add_argument("--early_stopping", type=int, default=3)
    p.add_argument("--continue_training", action="store_true")
    p.add_argument("--force_GPU", type=str, default="")
    p.add_argument("--eval_prob", type=float, default=1.0)
    p.add_argument("--wait_for", type=str, default="")
    p.add_argument("--dice_weight", type=str, default="uniform")
    return p


def entry_func(args=None):
    a = get_argparser().parse_args(args)
    base_dir = os.path.abspath(a.project_dir)
    if a.force_GPU:
        os.environ["CUDA_VISIBLE_DEVICES"] = a.force_GPU
    import torch
    from .. import distributed as D
    from .. import models
    from ..hyperparameters import YAMLHParams
    from ..image import ImagePairLoader
    from ..models import FusionModel
    from ..sequences import IsotrophicLiveViewSequence2D
    from ..utils.fusion import predict_and_map
    from ..utils.utils import get_best_model

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    D.init_from_env(device=torch.device("cuda", local))
    hp = YAMLHParams(os.path.join(base_dir, "train_hparams.yaml"))
    build, fit = dict(hp["build"]), dict(hp["fit"])
    views = np.load(os.path.join(base_dir, "views.npz"))["arr_0"]
    weights = get_best_model(os.path.join(base_dir, "model"))
    stem = os.path.splitext(os.path.basename(weights))[0]
    fdir = os.path.join(base_dir, "model", "fusion_weights")
    os.makedirs(fdir, exist_ok=True)
    fpath = os.path.join(fdir, "%s_fusion_weights.npz" % stem)
    if os.path.exists(fpath) and not (a.overwrite or a.continue_training):
        raise OSError("Fusion weights exist at %s (use --overwrite)" % fpath)
    model = models.__dict__[build["model_class_name"]](max_batch=32, training=False, **build)
    model.load_weights(weights)
    fm = FusionModel(n_inputs=len(views), n_classes=build["n_classes"], weight=a.dice_weight)
    if a.continue_training and os.path.exists(fpath):
        fm.load_weights(fpath)
    bg_value = hp.get_from_anywhere("bg_value") or 0.0
    loaders = [ImagePairLoader(bg_value=bg_value, **hp[k]) for k in ("train_data", "val_data")
               if hp[k].get("base_dir")]
    pairs = [(ld, i) for ld in loaders for i in range(len(ld))]
    pairs = D.shard(pairs)  # every rank builds the points of its own volumes
    rng = np.random.RandomState(0)
    order = rng.permutation(len(pairs))
    # round-robin shards differ by at most one volume: every rank runs the round count of the largest shard (a rank
    # whose shard is exhausted contributes an empty point set) so the collectives inside fit / evaluate stay paired
    n_pairs_max = torch.tensor([len(pairs)], device=model.device)
    if D.world_size() > 1:
        torch.distributed.all_reduce(n_pairs_max, op=torch.distributed.ReduceOp.MAX)
    for r0 in range(0, int(n_pairs_max.item()), a.images_per_round):
        Xs, ys = [], []
        for j in order[r0:r0 + a.images_per_round]:
            ld, i = pairs[j]
            image = ld.get(i)
            seq = IsotrophicLiveViewSequence2D([image], views=views, sample_dim=build["dim"],
                                               real_space_span=fit["real_space_span"],
                                               n_classes=build["n_classes"], is_validation=True)
            n_vox = int(np.prod(image.shape[:3]))
            X = torch.empty(n_vox, len(views), build["n_classes"], dtype=torch.float32, device=model.device)
            for k, v in enumerate(views):
                X[:, k, :] = predict_and_map(model, seq, image, v)
            Xs.append(X)
            ys.append(torch.as_tensor(image.labels.reshape(-1)).to(model.device))
            ld.unload(i)
        if Xs:
            X, y = torch.cat(Xs), torch.cat(ys)
        else:  # a rank without images in this round still takes part in every collective
            X = torch.zeros(0, len(views), build["n_classes"], dtype=torch.float32, device=model.device)
            y = torch.zeros(0, dtype=torch.uint8, device=model.device)
        perm = torch.randperm(X.shape[0], device=X.device)
        n_val = int(0.2 * X.shape[0])            # validation_split=0.2 of the reference's fit (train_fusion.py:205)
        tr, va = perm[n_val:], perm[:n_val]
        # every rank must run the same number of batches per epoch and of epochs: agree on the maximum
        steps = torch.tensor([(int(tr.numel()) + a.batch_size - 1) // a.batch_size], device=model.device)
        if D.world_size() > 1:
            torch.distributed.all_reduce(steps, op=torch.distributed.ReduceOp.MAX)
        steps = max(int(steps.item()), 1)
        best, wait = np.inf, 0
        for ep in range(a.epochs):
            fm.fit(X, y, batch_size=a.batch_size, epochs=1, verbose=0, index=tr, steps_per_epoch=steps)
            # val_loss drives early stopping (:199-203); evaluate() all-reduces, so a rank with no points still calls it
            loss = fm.evaluate(X[va], y[va]) if (n_val or D.world_size() > 1) else float("nan")
            if loss < best - 1e-6:
                best, wait = loss, 0
            else:
                wait += 1
                if wait >= a.early_stopping:
                    break
        if D.rank() == 0:
            print("round %d: fusion loss %.5f\nW=\n%s\nb=%s" % (r0 // a.images_per_round, best, *fm.get_weights()))
    if D.rank() == 0:
        fm.save_weights(fpath)
