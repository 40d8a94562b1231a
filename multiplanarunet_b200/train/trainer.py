"""The epoch loop `model.fit` runs for the reference (mpunet/train/trainer.py:246-257 hands it to Keras):
on_train_begin -> per epoch [on_epoch_begin, `steps_per_epoch` train steps, on_epoch_end(epoch, logs)] ->
on_train_end, stopping when a callback sets `model.stop_training`.  Pure host logic around the device train step."""
import numpy as np


def fit_loop(model, batches, steps_per_epoch, epochs, callbacks=None, initial_epoch=0, train_on_batch=None,
             verbose=1, logger=None, sync_stop=None):
    """batches: iterable of (x, y, w); train_on_batch: callable(x, y, w) -> mean loss (default: the model's);
    sync_stop: optional callable(bool) -> bool agreeing on the stop flag across ranks.
    Returns {"loss": [...], <every scalar a callback logged>: [...]} like Keras' History.history."""
    callbacks = list(callbacks or [])
    # the model's asynchronous step (loss stays on the device) lets the host queue the next step while this one
    # runs; the losses of an epoch are read back once, at its end
    step = train_on_batch or getattr(model, "train_on_batch_async", None) or model.train_on_batch
    log = logger or getattr(model, "logger", print)
    history = {}
    for cb in callbacks:
        if hasattr(cb, "set_model"):
            cb.set_model(model)
    model.stop_training = False
    it = iter(batches)
    for cb in callbacks:
        cb.on_train_begin()
    try:
        for epoch in range(initial_epoch, epochs):
            for cb in callbacks:
                cb.on_epoch_begin(epoch)
            losses = []
            for _ in range(steps_per_epoch or 1):
                bx, by, bw = next(it)
                losses.append(step(bx, by, bw))
            logs = {"loss": float(np.mean([float(v.item()) if hasattr(v, "item") else float(v) for v in losses]))}
            for cb in callbacks:
                cb.on_epoch_end(epoch, logs)
            for k, v in logs.items():
                if np.isscalar(v):
                    history.setdefault(k, []).append(float(v))
            if verbose:
                log("Epoch %d/%d - %s" % (epoch + 1, epochs, " - ".join(
                    "%s: %.5g" % kv for kv in sorted(logs.items()) if np.isscalar(kv[1]))))
            if sync_stop is not None:
                model.stop_training = bool(sync_stop(bool(model.stop_training)))
            if model.stop_training:
                break
    finally:
        for cb in callbacks:
            cb.on_train_end()
    return history
