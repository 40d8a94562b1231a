"""Epoch-level callbacks of `mp train`, selected by class name from train_hparams.yaml
(mpunet/bin/defaults/MultiPlanar/train_hparams.yaml:7-44, mpunet/callbacks/funcs.py:5-58).

The reference takes ReduceLROnPlateau / EarlyStopping / CSVLogger / TensorBoard from tf.keras (2.3) and adds its own
ModelCheckPointClean, TrainTimer, FGBatchBalancer, DelayedCallback, DividerLine (mpunet/callbacks/callbacks.py,
mcp_clean.py).  There is no Keras here: the classes below restate the documented Keras behaviour for the keyword
arguments the reference's YAML uses; they are pure host logic (no device work) driven by the training loop in
bin/train.py through on_train_begin / on_epoch_begin / on_epoch_end, with `model.optimizer.lr` and
`model.stop_training` as the only model attributes touched."""
import csv
import os
import warnings

import numpy as np


class Callback(object):
    def __init__(self):
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass


def _monitor_op(mode, monitor, min_delta=0.0):
    """(op, initial best) of Keras' monitor handling: 'auto' means max for accuracy-like names."""
    if mode not in ("auto", "min", "max"):
        warnings.warn("mode %s is unknown, fallback to auto mode." % mode, RuntimeWarning)
        mode = "auto"
    if mode == "min" or (mode == "auto" and "acc" not in monitor and not monitor.startswith("fmeasure")):
        return (lambda a, b: np.less(a, b - min_delta)), np.inf
    return (lambda a, b: np.greater(a, b + min_delta)), -np.inf


class ReduceLROnPlateau(Callback):
    """keras.callbacks.ReduceLROnPlateau (monitor, factor, patience, mode, min_delta, cooldown, min_lr, verbose)."""

    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, mode="auto", min_delta=1e-4,
                 cooldown=0, min_lr=0, **kwargs):
        super().__init__()
        if factor >= 1.0:
            raise ValueError("ReduceLROnPlateau does not support a factor >= 1.0.")
        self.monitor, self.factor, self.patience, self.verbose = monitor, factor, patience, verbose
        self.mode, self.min_delta, self.cooldown, self.min_lr = mode, min_delta, cooldown, min_lr
        self._reset()

    def _reset(self):
        self.monitor_op, self.best = _monitor_op(self.mode, self.monitor, self.min_delta)
        self.cooldown_counter = 0
        self.wait = 0

    def on_train_begin(self, logs=None):
        self._reset()

    def in_cooldown(self):
        return self.cooldown_counter > 0

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        logs["lr"] = float(self.model.optimizer.lr)
        current = logs.get(self.monitor)
        if current is None:
            warnings.warn("Reduce LR on plateau conditioned on metric `%s` which is not available. Available "
                          "metrics are: %s" % (self.monitor, ",".join(list(logs.keys()))), RuntimeWarning)
            return
        if self.in_cooldown():
            self.cooldown_counter -= 1
            self.wait = 0
        if self.monitor_op(current, self.best):
            self.best = current
            self.wait = 0
        elif not self.in_cooldown():
            self.wait += 1
            if self.wait >= self.patience:
                old_lr = float(self.model.optimizer.lr)
                if old_lr > self.min_lr:
                    new_lr = max(old_lr * self.factor, self.min_lr)
                    self.model.optimizer.lr = new_lr
                    if self.verbose > 0:
                        print("\nEpoch %05d: ReduceLROnPlateau reducing learning rate to %s." % (epoch + 1, new_lr))
                    self.cooldown_counter = self.cooldown
                    self.wait = 0


class EarlyStopping(Callback):
    """keras.callbacks.EarlyStopping (monitor, min_delta, patience, mode, baseline, verbose)."""

    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto", baseline=None,
                 restore_best_weights=False):
        super().__init__()
        if restore_best_weights:
            raise NotImplementedError("restore_best_weights is not used by the reference's presets")
        self.monitor, self.patience, self.verbose, self.baseline = monitor, patience, verbose, baseline
        self.min_delta = abs(min_delta)
        op, _ = _monitor_op(mode, monitor)
        self._greater = bool(op(1, 0))
        self.monitor_op = np.greater if self._greater else np.less
        self.min_delta *= 1 if self._greater else -1
        self.wait = 0
        self.stopped_epoch = 0
        self.best = None

    def on_train_begin(self, logs=None):
        self.wait = 0
        self.stopped_epoch = 0
        if self.baseline is not None:
            self.best = self.baseline
        else:
            self.best = -np.inf if self._greater else np.inf

    def on_epoch_end(self, epoch, logs=None):
        current = (logs or {}).get(self.monitor)
        if current is None:
            warnings.warn("Early stopping conditioned on metric `%s` which is not available." % self.monitor,
                          RuntimeWarning)
            return
        if self.monitor_op(current - self.min_delta, self.best):
            self.best = current
            self.wait = 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.model.stop_training = True

    def on_train_end(self, logs=None):
        if self.stopped_epoch > 0 and self.verbose > 0:
            print("Epoch %05d: early stopping" % (self.stopped_epoch + 1))


class ModelCheckPointClean(Callback):
    """mpunet/callbacks/mcp_clean.py: ModelCheckpoint that removes the previous best file when the formatted
    file name changes.  `filepath` keeps the reference's pattern; the `.h5` suffix becomes `.npz`
    (weights are stored by Keras layer name in numpy archives here)."""

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", period=1, org_model=None, **kwargs):
        super().__init__()
        self.filepath = filepath[:-3] + ".npz" if filepath.endswith(".h5") else filepath
        self.monitor, self.verbose, self.save_best_only = monitor, verbose, save_best_only
        self.save_weights_only, self.period = save_weights_only, period
        self.epochs_since_last_save = 0
        self.org_model = org_model
        self.last_file = None
        op, self.best = _monitor_op(mode, monitor)
        self.monitor_op = np.greater if op(1, 0) else np.less

    def _model(self):
        return self.org_model if self.org_model is not None else self.model

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        # mcp_clean.py:33 formats with the 0-based epoch index of on_epoch_end
        filepath = self.filepath.format(epoch=epoch, **logs)
        folder = os.path.split(os.path.abspath(filepath))[0]
        if not os.path.exists(folder):
            os.mkdir(folder)
        if self.save_best_only:
            current = logs.get(self.monitor)
            if current is None:
                warnings.warn("Can save best model only with %s available, skipping." % self.monitor,
                              RuntimeWarning)
                return
            if self.monitor_op(current, self.best):
                if self.verbose > 0:
                    print("Epoch %05d: %s improved from %0.5f to %0.5f, saving model to %s"
                          % (epoch, self.monitor, self.best, current, filepath))
                self.best = current
                if self.last_file and os.path.exists(self.last_file):
                    os.remove(self.last_file)
                self.last_file = filepath
                self._model().save_weights(filepath)
            elif self.verbose > 0:
                print("Epoch %05d: %s did not improve" % (epoch, self.monitor))
        else:
            if self.verbose > 0:
                print("Epoch %05d: saving model to %s" % (epoch, filepath))
            self._model().save_weights(filepath)


class CSVLogger(Callback):
    """keras.callbacks.CSVLogger (filename, separator, append): one row per epoch, columns = 'epoch' + the sorted
    keys of the first epoch's logs, missing later values written as 'NA'."""

    def __init__(self, filename, separator=",", append=False):
        super().__init__()
        self.filename, self.sep, self.append = filename, separator, append
        self.keys = None
        self.append_header = True
        self.file = None
        self.writer = None

    def on_train_begin(self, logs=None):
        if self.append and os.path.exists(self.filename):
            with open(self.filename, "r") as f:
                self.append_header = not bool(len(f.readline()))
        folder = os.path.dirname(os.path.abspath(self.filename))
        os.makedirs(folder, exist_ok=True)
        self.file = open(self.filename, "a" if self.append else "w", newline="")

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}

        def fmt(v):
            if isinstance(v, np.ndarray) and v.ndim > 0:
                return '"[%s]"' % ", ".join(map(str, v))
            return v
        if self.keys is None:
            self.keys = sorted(logs.keys())
        if self.writer is None:
            self.writer = csv.DictWriter(self.file, fieldnames=["epoch"] + self.keys, delimiter=self.sep)
            if self.append_header:
                self.writer.writeheader()
        row = {"epoch": epoch}
        row.update((k, fmt(logs[k]) if k in logs else "NA") for k in self.keys)
        self.writer.writerow(row)
        self.file.flush()

    def on_train_end(self, logs=None):
        if self.file:
            self.file.close()
        self.file = None
        self.writer = None


class TensorBoard(Callback):
    """TensorBoard event files are outside this build's scope (no arithmetic on the hot path);
    the class exists so that the default YAML's `tb` entry initialises, and says so once."""

    def __init__(self, logger=None, **kwargs):
        super().__init__()
        (logger or print)("[NOTE] TensorBoard logging is not available on the B200 path - callback is a no-op")


class DividerLine(Callback):
    """mpunet/callbacks/callbacks.py:15-29."""

    def __init__(self, logger=None):
        super().__init__()
        self.logger = logger or print

    def on_epoch_end(self, epoch, logs=None):
        self.logger("-" * 45 + "\n")


class DelayedCallback(object):
    """Holds another callback back until epoch `start_from` (counted from 1, like the YAML's `start_from`): its
    on_epoch_end is forwarded only from then on; every other attribute resolves to the wrapped object.
    Behaviour of the reference's wrapper of the same name (mpunet/callbacks/callbacks.py:88-115)."""

    def __init__(self, callback, start_from=0, logger=None):
        self.callback = callback
        self.start_from = int(start_from)
        self.logger = logger or print

    def __getattr__(self, name):
        # only called for attributes this wrapper does not define itself
        return getattr(self.callback, name)

    def is_active(self, epoch):
        return epoch + 1 >= self.start_from

    def on_epoch_end(self, epoch, logs=None):
        if self.is_active(epoch):
            return self.callback.on_epoch_end(epoch, logs=logs)
        self.logger("[%s] waiting: epoch %d < start_from %d" % (type(self.callback).__name__, epoch + 1,
                                                                self.start_from))


class TrainTimer(Callback):
    """Wall-clock bookkeeping per epoch: writes `epoch_minutes` and `train_hours` into the epoch logs (so CSVLogger
    records them) and stops training once `max_minutes` of total time are used up.  Log keys and the stop rule follow
    the reference's TrainTimer (mpunet/callbacks/callbacks.py:118-163)."""

    def __init__(self, logger=None, max_minutes=None, verbose=1):
        super().__init__()
        self.logger = logger or print
        self.max_minutes = int(max_minutes) if max_minutes else None
        self.verbose = bool(verbose)
        self._t_train = None
        self._t_epoch = None

    @staticmethod
    def _now():
        import time
        return time.monotonic()

    def on_train_begin(self, logs=None):
        self._t_train = self._now()

    def on_epoch_begin(self, epoch, logs=None):
        self._t_epoch = self._now()

    def on_epoch_end(self, epoch, logs=None):
        now = self._now()
        minutes = (now - (self._t_epoch if self._t_epoch is not None else now)) / 60.0
        hours = (now - (self._t_train if self._t_train is not None else now)) / 3600.0
        self._t_epoch = now
        if logs is not None:
            logs["epoch_minutes"] = round(minutes, 4)
            logs["train_hours"] = round(hours, 4)
        if self.verbose:
            self.logger("[TrainTimer] epoch took %.2f min, training so far %.2f h" % (minutes, hours))
        if self.max_minutes and hours * 60.0 > self.max_minutes:
            self.logger("[TrainTimer] time budget of %d min exhausted after %.1f min: stopping" % (self.max_minutes,
                                                                                                   hours * 60.0))
            self.model.stop_training = True


class FGBatchBalancer(Callback):
    """Adapts the fraction of slices per batch that must contain foreground to how much foreground the model still
    misses: fg_batch_fraction := max(0.01, 1 - val_recall) on the training (and optionally validation) sequence after
    every epoch.  Needs the Validation callback to have run first in the same epoch; switches itself off when the
    logs carry no `val_recall`.  Rule of the reference's FGBatchBalancer (mpunet/callbacks/callbacks.py:166-209)."""

    MIN_FRACTION = 0.01

    def __init__(self, train_data, val_data=None, logger=None):
        super().__init__()
        self.sequences = {"train": train_data, "val": val_data}
        self.logger = logger or print
        self.active = True

    def on_epoch_end(self, epoch, logs=None):
        if not self.active:
            return
        recall = None if logs is None else logs.get("val_recall")
        if recall is None:
            self.active = False
            self.logger("[FGBatchBalancer] the epoch logs hold no val_recall (is this callback listed after "
                        "Validation?) - switching off")
            return
        fraction = max(self.MIN_FRACTION, 1.0 - float(recall))
        for name, seq in self.sequences.items():
            if seq is None:
                continue
            seq.fg_batch_fraction = fraction
            self.logger("[FGBatchBalancer] %s: fg_batch_fraction = %.4f (%d of %d slices per batch)"
                        % (name, fraction, seq.n_fg_slices, seq.batch_size))
