"""Project-directory helpers mirroring mpunet/utils/utils.py (best/last model selection :88-130,
output-bias initialisation :205-242, folder creation)."""
import glob
import os
import re

import numpy as np

WEIGHT_EXTS = (".npz", ".h5")


def _models_in(model_dir):
    out = []
    for ext in WEIGHT_EXTS:
        out += glob.glob(os.path.join(model_dir, "@epoch*" + ext))
    return out


def get_best_model(model_dir):
    """Best checkpoint by the reference's rule (utils.py:88-110): the first of the name patterns `@epoch*val_dice*`
    (highest score), `@epoch*val_loss*` (lowest), `@epoch*dice*` (highest), `@epoch*loss*` (lowest) that matches any
    file decides; the score is the first decimal number in the file name; otherwise `model_weights.*`."""
    if len(os.listdir(model_dir)) == 0:
        raise OSError("Model dir {} is empty.".format(model_dir))
    models = _models_in(model_dir)
    for key, pick in (("val_dice", np.argmax), ("val_loss", np.argmin), ("dice", np.argmax), ("loss", np.argmin)):
        cand = [m for m in models if key in os.path.basename(m)]
        if cand:
            scores = [float(re.findall(r"(\d+[.]\d+)", os.path.basename(m))[0]) for m in cand]
            return os.path.abspath(cand[int(pick(np.array(scores)))])
    for ext in WEIGHT_EXTS:
        p = os.path.join(model_dir, "model_weights" + ext)
        if os.path.exists(p):
            return os.path.abspath(p)
    raise OSError("Did not find any model files in directory {}".format(model_dir))


def get_last_model(model_dir):
    """(path, epoch) of the checkpoint with the largest epoch number (utils.py:113-130)."""
    models = _models_in(model_dir)
    if not models:
        for ext in WEIGHT_EXTS:
            p = os.path.join(model_dir, "model_weights" + ext)
            if os.path.exists(p):
                return p, 0
        return None, 0
    epochs = [int(re.findall(r"@epoch_(\d+)", os.path.basename(m))[0]) for m in models]
    i = int(np.argmax(epochs))
    return os.path.abspath(models[i]), epochs[i]


def _read_csv(csv_file):
    import csv
    with open(csv_file, newline="") as f:
        rows = list(csv.reader(f))
    return (rows[0], rows[1:]) if rows else ([], [])


def get_last_epoch(csv_file):
    """Last value of the `epoch` column of logs/training.csv, 0 if there is none (utils.py:171-177)."""
    if not os.path.exists(csv_file):
        return 0
    header, rows = _read_csv(csv_file)
    if "epoch" not in header or not rows:
        return 0
    return int(float(rows[-1][header.index("epoch")]))


def get_lr_at_epoch(epoch, log_dir):
    """(lr, column name) logged for `epoch` in logs/training.csv, (None, None) if unavailable (utils.py:133-147)."""
    log_path = os.path.join(log_dir, "training.csv")
    if not os.path.exists(log_path):
        print("No training.csv file found at %s. Continuing with default learning rate found in parameter file."
              % log_dir)
        return None, None
    header, rows = _read_csv(log_path)
    for name in ("lr", "LR", "learning_rate", "LearningRate"):
        if name in header:
            if int(epoch) >= len(rows):
                return None, None
            return float(rows[int(epoch)][header.index(name)]), name
    return None, None


def clear_csv_after_epoch(epoch, csv_file):
    """Keeps the last run in the file (from its last `epoch == 0` row) up to and including `epoch`
    (utils.py:150-168)."""
    if not os.path.exists(csv_file):
        return
    header, rows = _read_csv(csv_file)
    if not header:
        os.remove(csv_file)
        return
    if "epoch" in header:
        col = header.index("epoch")
        zeros = [i for i, r in enumerate(rows) if int(float(r[col])) == 0]
        if zeros:
            rows = rows[zeros[-1]:]
    rows = rows[:epoch + 1]
    import csv
    with open(csv_file, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(header)
        w.writerows(rows)


def create_folders(folders, create_deep=False):
    folders = [folders] if isinstance(folders, str) else folders
    for f in folders:
        os.makedirs(f, exist_ok=True) if create_deep else (os.path.isdir(f) or os.mkdir(f))


def set_bias_weights(layer, data_queue=None, class_counts=None, logger=None):
    """Initialise the softmax layer's bias so that a zero input gives the class frequencies
    (utils.py:205-242): b = log(freq * sum(exp(freq))) normalised to unit length.  `data_queue`: iterable of
    images with `.labels` (counted when `class_counts` is None)."""
    if getattr(layer.activation, "__name__", None) != "softmax":
        raise ValueError("Setting output layer bias currently only supported with softmax activation functions. "
                         "Output layer has '%s'" % getattr(layer.activation, "__name__", layer.activation))
    weights = layer.get_weights()
    if len(weights) != 2:
        raise ValueError("Output layer does not have bias weights.")
    bias_shape = weights[-1].shape
    n_classes = weights[-1].size
    if class_counts is None:
        class_counts = np.zeros(shape=[n_classes], dtype=np.int64)
        images = list(data_queue)
        (logger or print)("OBS: Estimating class counts from {} images".format(len(images)))
        for image in images:
            class_counts += np.bincount(np.asarray(image.labels).ravel(), minlength=n_classes)[:n_classes]
    counts = np.asarray(class_counts, dtype=np.float64)
    freq = counts / np.sum(counts)
    bias = np.log(freq * np.sum(np.exp(freq)))
    bias /= np.linalg.norm(bias)
    weights[-1] = bias.reshape(bias_shape).astype(np.float32)
    layer.set_weights(weights)
    (logger or print)("Setting bias weights on output layer to:\n%s" % bias)
    return bias


def set_bias_weights_on_all_outputs(model, data_queue, hparams, logger=None):
    """utils.py:179-202: the last layer that has an activation is the output layer."""
    layer = None
    for lay in model.layers[::-1]:
        if hasattr(lay, "activation"):
            layer = lay
            break
    class_counts = hparams.get("class_counts") if hasattr(hparams, "get") else None
    return set_bias_weights(layer=layer, data_queue=data_queue, class_counts=class_counts, logger=logger)


def pred_to_class(tensor, img_dims=3, threshold=0.5, has_batch_dim=False):
    """utils.py:311-328: soft-max volume -> uint8 class map."""
    tensor = np.asarray(tensor)
    tensor_dim = img_dims + int(has_batch_dim)
    dims = len(tensor.shape)
    if dims == tensor_dim:
        return tensor
    if tensor.shape[-1] == 1:
        return (tensor.squeeze(-1) > threshold).astype(np.uint8)
    return tensor.argmax(-1).astype(np.uint8)
