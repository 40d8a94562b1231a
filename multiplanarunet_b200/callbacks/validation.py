"""Epoch-end validation (mpunet/callbacks/validation.py:14-300): samples `steps` validation batches, predicts,
and accumulates TP / relevant / selected per class over the whole epoch - here with `mpu_label_counts` on the
device instead of three np.bincount calls per batch on a counting thread - then derives per-class precision,
recall and dice exactly like `_compute_dice` and writes their nan-means to the logs as val_precision / val_recall /
val_dice."""
import numpy as np

from ..evaluate import compute_dice, label_counts
from .callbacks import Callback


class Validation(Callback):
    def __init__(self, val_sequence, steps, logger=None, verbose=True, ignore_class_zero=True):
        super().__init__()
        self.logger = logger or print
        self.data = val_sequence
        self.steps = steps
        self.verbose = verbose
        self.ignore_bg = ignore_class_zero
        self.print_round = 3
        self.log_round = 4
        self.n_classes = int(self.data.n_classes)

    def evalaute(self):  # (sic) name kept from the reference
        counts = None
        for _ in range(self.steps):
            x, y, _w = self.data.sample_batch_device()
            probs = self.model.predict_on_batch(x, as_numpy=False)
            counts = label_counts(y, probs, self.n_classes, counts=counts)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(counts)  # identical metrics (hence LR / stopping decisions) on every rank
        c = counts.cpu().numpy().astype(np.uint64)
        # validation.py:213-215 calls _compute_dice(tp=TPs, sel=relevant, rel=selected): the names are swapped
        # there, so "precision" is TP/relevant and "recall" is TP/selected - kept, FGBatchBalancer reads it.
        precisions, recalls, dices = compute_dice(tp=c[0], sel=c[1], rel=c[2])
        if self.ignore_bg:
            precisions[0] = np.nan
            recalls[0] = np.nan
            dices[0] = np.nan
        return {"dice": dices, "recall": recalls, "precision": precisions}

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        cw = self.evalaute()
        for name, values in cw.items():
            logs["val_%s" % name] = float(np.nanmean(values)) if np.any(~np.isnan(values)) else float("nan")
        if self.verbose:
            self.logger("Validation Results for epoch %i" % epoch)
            header = "%-10s" % "" + "".join("%10s" % k for k in cw)
            self.logger(header)
            self.logger("%-10s" % "mean" + "".join("%10.*f" % (self.print_round, logs["val_%s" % k]) for k in cw))
            for cls in range(self.n_classes):
                self.logger("%-10s" % ("cls %i" % cls) + "".join(
                    "%10s" % ("-" if np.isnan(cw[k][cls]) else "%.*f" % (self.print_round, cw[k][cls]))
                    for k in cw))
