"""`mp train` on the B200 engine (mirror of mpunet/bin/train.py:18-416 for MultiPlanar projects).

Keeps the reference's flags, project-dir layout (train_hparams.yaml, views.npz, model/, logs/) and
training semantics (Adam, sparse CE x sample weight, FG-balanced random oblique slices, per-epoch
validation dice driving ReduceLROnPlateau / best-checkpoint / early stopping with the YAML presets);
the Keras fit loop is replaced by a thin step loop over the device sampler and the CUDA train step.
Launch under torchrun for multi-GPU (one process per GPU, NCCL gradient all-reduce)."""
import math
import os
import shutil
from argparse import ArgumentParser

import numpy as np


def get_argparser():
    p = ArgumentParser(description='Fit a mpunet model defined in a project folder. '
                                   'Invoke "init_project" to start a new project.')
    p.add_argument("--project_dir", type=str, default="./")
    p.add_argument("--num_GPUs", type=int, default=1)
    p.add_argument("--force_GPU", type=str, default="")
    p.add_argument("--continue_training", action="store_true")
    p.add_argument("--overwrite", action="store_true")
    p.add_argument("--just_one", action="store_true")
    p.add_argument("--no_val", action="store_true")
    p.add_argument("--no_images", action="store_true")
    p.add_argument("--debug", action="store_true")
    p.add_argument("--wait_for", type=str, default="")
    p.add_argument("--train_images_per_epoch", type=int, default=2500)
    p.add_argument("--val_images_per_epoch", type=int, default=3500)
    p.add_argument("--max_loaded_images", type=int, default=None)
    p.add_argument("--epochs", type=int, default=None)
    p.add_argument("--num_access", type=int, default=50)
    return p


def validate_project_dir(project_dir):
    if not os.path.exists(project_dir) or not os.path.exists(os.path.join(project_dir, "train_hparams.yaml")):
        raise RuntimeError("The script was launched from directory:\n'%s'\n... but this is not a valid project "
                           "folder.\n\n* Make sure to launch the script from within a MultiPlanarNet project "
                           "directory\n* Make sure that the directory contains a 'train_hparams.yaml' file."
                           % project_dir)


def remove_previous_session(project_folder):
    for p in ("images", "logs", "model", "tensorboard", "views.npz", "views.png"):
        p = os.path.join(project_folder, p)
        if os.path.isdir(p):
            shutil.rmtree(p)
        elif os.path.exists(p):
            os.remove(p)


def load_or_create_views(project_dir, n_views, continue_training):
    from ..interpolation import sample_random_views_with_angle_restriction
    path = os.path.join(project_dir, "views.npz")
    if continue_training and os.path.exists(path):
        return np.load(path)["arr_0"]
    views = sample_random_views_with_angle_restriction(n_views, 60)
    np.savez(path, views)
    return views


def run(project_dir, args):
    import torch
    from .. import distributed as D
    from ..hyperparameters import YAMLHParams
    from ..image import Auditor, ImagePairLoader
    from .. import models
    from ..sequences import IsotrophicLiveViewSequence2D
    from ..utils.utils import get_last_model

    hp = YAMLHParams(os.path.join(project_dir, "train_hparams.yaml"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    world = D.init_from_env(device=torch.device("cuda", local))
    rank = D.rank()
    log = print if rank == 0 else (lambda *a, **k: None)

    bg_value = hp.get_from_anywhere("bg_value") or 0.0
    train = ImagePairLoader(bg_value=bg_value, **hp["train_data"])
    val = None if args.no_val else ImagePairLoader(bg_value=bg_value, **hp["val_data"])
    if len(train) == 0:
        raise OSError("No training images found under %s" % hp["train_data"]["base_dir"])
    if args.just_one:
        train.image_paths = train.image_paths[:1]
        if val:
            val.image_paths = val.image_paths[:1]
    if rank == 0:
        paths = train.image_paths + (val.image_paths if val else [])
        labs = [train.label_path_for(p) for p in train.image_paths]
        Auditor(paths, [l for l in labs if l], hparams=hp).fill(hp, "2d")
    if world > 1:
        torch.distributed.barrier()
        hp = YAMLHParams(os.path.join(project_dir, "train_hparams.yaml"))
    build, fit = dict(hp["build"]), dict(hp["fit"])
    views = load_or_create_views(project_dir, fit["views"], args.continue_training) if rank == 0 else None
    if world > 1:
        torch.distributed.barrier()
    if views is None:
        views = np.load(os.path.join(project_dir, "views.npz"))["arr_0"]

    # volumes shard across ranks (one or more resident volumes per GPU); every rank keeps >= 1
    my_train = D.shard(list(range(len(train)))) or [rank % len(train)]
    train_images = [train.get(i) for i in my_train]
    val_images = [val.get(i) for i in range(len(val))] if val else []
    bs = int(fit["batch_size"])
    common = dict(views=views, sample_dim=build["dim"], real_space_span=fit["real_space_span"],
                  n_classes=build["n_classes"], batch_size=bs, fg_batch_fraction=fit.get("fg_batch_fraction", 0.5))
    aug_list = []
    if fit.get("augmenters"):
        # sequences/utils.py:38-47: classes looked up by name, kwargs from the YAML
        from .. import augmentation
        for aug in fit["augmenters"]:
            if aug["cls_name"] not in augmentation.__dict__:
                raise NotImplementedError("augmenter %r is not available on the B200 path (2D pipeline: "
                                          "Elastic2D)" % aug["cls_name"])
            aug_list.append(augmentation.__dict__[aug["cls_name"]](**aug["kwargs"]))
        log("Using on-the-fly augmenters: %s" % aug_list)
    tr_seq = IsotrophicLiveViewSequence2D(train_images, noise_sd=fit.get("noise_sd", 0.1),
                                          list_of_augmenters=aug_list, **common)
    va_seq = IsotrophicLiveViewSequence2D(val_images, is_validation=True, **common) if val_images else None

    cls = models.__dict__[build["model_class_name"]]
    model = cls(max_batch=bs, training=True, logger=log, **{k: v for k, v in build.items() if k != "logger"})
    okw = fit.get("optimizer_kwargs", {})
    model.optimizer.lr = float(okw.get("lr", 5e-5))
    model.optimizer.beta_1, model.optimizer.beta_2 = float(okw.get("beta_1", 0.9)), float(okw.get("beta_2", 0.999))
    model.optimizer.epsilon = float(okw.get("epsilon", 1e-7))
    os.makedirs(os.path.join(project_dir, "model"), exist_ok=True)
    os.makedirs(os.path.join(project_dir, "logs"), exist_ok=True)
    init_epoch = 0
    if args.continue_training:
        # models/model_init.py:25-51: last checkpoint, its epoch (or the CSV's last epoch), the LR logged there
        from ..utils.utils import clear_csv_after_epoch, get_last_epoch, get_lr_at_epoch
        last, epoch = get_last_model(os.path.join(project_dir, "model"))
        if last:
            model.load_weights(last)
        csv_file = os.path.join(project_dir, "logs", "training.csv")
        if epoch == 0:
            epoch = get_last_epoch(csv_file)
        elif rank == 0:
            clear_csv_after_epoch(epoch, csv_file)
        if world > 1:
            torch.distributed.barrier()
        init_epoch = epoch + 1
        lr, _ = get_lr_at_epoch(epoch, os.path.join(project_dir, "logs"))
        if lr:
            model.optimizer.lr = lr
        log("[NOTICE] Training continues from:\nModel: %s\nEpoch: %i\nLR:    %s"
            % (os.path.split(last)[-1] if last else "<No model found>", epoch, lr))
    elif build.get("biased_output_layer"):
        # bin/train.py:294-299: output bias from the training set's class frequencies (summed over all ranks'
        # shards so that every replica starts from identical weights)
        from ..utils.utils import set_bias_weights_on_all_outputs
        k = int(build["n_classes"])
        counts = np.zeros(k, dtype=np.int64)
        for image in train_images:
            if image.labels is not None:
                counts += np.bincount(np.asarray(image.labels).ravel(), minlength=k)[:k]
        if world > 1:
            t = torch.as_tensor(counts, device="cuda")
            torch.distributed.all_reduce(t)
            counts = t.cpu().numpy()
        if counts.sum() > 0 and np.all(counts > 0):
            set_bias_weights_on_all_outputs(model, train_images, {"class_counts": counts}, log)
        else:
            log("[NOTE] output bias not initialised from class frequencies (classes missing: %s)" % counts)
    dp = D.DataParallel(model)

    n_epochs = args.epochs or int(fit["n_epochs"])
    steps = int(math.ceil(args.train_images_per_epoch / (bs * world)))

    # ---- callbacks (train/trainer.py:180-232): Validation first, then the YAML list, FGBatchBalancer, divider
    from ..callbacks import (DividerLine, FGBatchBalancer, Validation, init_callback_objects,
                             remove_validation_callbacks)
    cb_descr = [dict(c) for c in fit.get("callbacks", []) if isinstance(c, dict)]
    if rank != 0:  # files are written by rank 0 only
        cb_descr = [c for c in cb_descr if c["class_name"] not in ("ModelCheckPointClean", "CSVLogger",
                                                                     "TensorBoard")]
    if va_seq is None:
        remove_validation_callbacks(cb_descr, log)
        callbacks = list(cb_descr)
    else:
        val_steps = max(1, int(math.ceil(args.val_images_per_epoch / bs)))
        callbacks = [Validation(va_seq, steps=val_steps, logger=log, verbose=bool(fit.get("verbose", True)),
                                ignore_class_zero=bool(fit.get("val_ignore_class_zero", True)))] + cb_descr
    callbacks.append(FGBatchBalancer(tr_seq, logger=log))
    callbacks.append(DividerLine(log))
    cwd = os.getcwd()
    os.chdir(project_dir)  # the YAML's callback paths are relative to the project (bin/train.py:398)
    callbacks, cb_dict = init_callback_objects(callbacks, log)
    if "CSVLogger" in cb_dict and args.continue_training:
        cb_dict["CSVLogger"].append = True

    def batches():
        while True:
            x, y, w = tr_seq.sample_batch_device()
            yield x, y, torch.as_tensor(w)

    def sync_stop(flag):  # every rank takes the same stop decision
        if world == 1:
            return flag
        t = torch.tensor([1.0 if flag else 0.0], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return bool(t.item() > 0)

    try:
        model.fit(batches(), steps_per_epoch=steps, epochs=n_epochs, callbacks=callbacks,
                  initial_epoch=init_epoch, verbose=1, train_on_batch=dp.train_on_batch, sync_stop=sync_stop)
    except KeyboardInterrupt:
        pass
    finally:
        os.chdir(cwd)
        if rank == 0:
            model.save_weights(os.path.join(project_dir, "model", "model_weights.npz"))


def entry_func(args=None):
    a = get_argparser().parse_args(args)
    project_dir = os.path.abspath(a.project_dir)
    validate_project_dir(project_dir)
    if a.overwrite and a.continue_training:
        raise ValueError("Cannot both continue training and overwrite the previous training session.")
    if not a.overwrite and not a.continue_training and os.path.exists(os.path.join(project_dir, "model")) and \
            os.listdir(os.path.join(project_dir, "model")):
        raise OSError("There seems to be a previous training session at '%s'. Use --overwrite or "
                      "--continue_training." % project_dir)
    if a.force_GPU:
        os.environ["CUDA_VISIBLE_DEVICES"] = a.force_GPU
    if a.overwrite:
        # under torchrun only rank 0 deletes, and nobody touches the project dir before it is done
        world = int(os.environ.get("WORLD_SIZE", "1"))
        marker = os.path.join(project_dir, ".overwrite_done_%s" % os.environ.get("TORCHELASTIC_RUN_ID", "single"))
        if int(os.environ.get("RANK", "0")) == 0:
            remove_previous_session(project_dir)
            if world > 1:
                open(marker, "w").close()
        elif world > 1:
            import time
            for _ in range(600):
                if os.path.exists(marker):
                    break
                time.sleep(0.1)
    run(project_dir, a)
    if a.overwrite and int(os.environ.get("RANK", "0")) == 0:
        try:
            os.remove(os.path.join(project_dir, ".overwrite_done_%s" % os.environ.get("TORCHELASTIC_RUN_ID", "single")))
        except OSError:
            pass
