"""Callback descriptors of train_hparams.yaml -> callback objects (behaviour of mpunet/callbacks/funcs.py:5-84).

A descriptor is a dict {"class_name", "kwargs"[, "start_from", "pass_logger", "nickname"]}; already constructed
callback objects pass through unchanged.  Classes are resolved by name in this package (the tf.keras classes the
reference falls back to - ReduceLROnPlateau, EarlyStopping, CSVLogger, TensorBoard - are restated in callbacks.py)."""
from .callbacks import DelayedCallback


def _resolve(cls_name):
    from . import callbacks as own
    from . import validation as val
    for module in (own, val):
        cls = getattr(module, cls_name, None)
        if cls is not None:
            return cls
    raise ValueError("No callback named %s" % cls_name)


def _build(descriptor, logger):
    """-> (object, class name, kwargs used, start_from)"""
    if not isinstance(descriptor, dict):
        return descriptor, type(descriptor).__name__, {"params": "?"}, 0
    kwargs = dict(descriptor.get("kwargs") or {})
    if descriptor.get("pass_logger"):
        kwargs["logger"] = logger
    name = descriptor["class_name"]
    return _resolve(name)(**kwargs), name, kwargs, descriptor.get("start_from")


def init_callback_objects(callbacks, logger):
    """Returns (list of callback objects in the given order, {class name: object})."""
    objects, by_name = [], {}
    for pos, descriptor in enumerate(callbacks, start=1):
        cb, name, kwargs, start_from = _build(descriptor, logger)
        if start_from:
            logger("OBS: '%s' activates at epoch %i" % (name, start_from))
            cb = DelayedCallback(callback=cb, start_from=start_from, logger=logger)
        objects.append(cb)
        by_name[name] = cb
        shown = ", ".join("%s=%s" % item for item in kwargs.items())
        logger("[%i] Using callback: %s(%s)" % (pos, type(cb).__name__, shown))
    return objects, by_name


def remove_validation_callbacks(callbacks, logger=None):
    """In place: drops every descriptor one of whose kwargs mentions 'val' (it needs validation data).  The
    reference pops while enumerating and so can skip the entry after a removed one (funcs.py:72-84); this
    implements the stated intent."""
    def needs_val(descriptor):
        return isinstance(descriptor, dict) and any(
            "val" in str(value).lower() for value in (descriptor.get("kwargs") or {}).values())
    kept = []
    for descriptor in callbacks:
        if needs_val(descriptor):
            if logger:
                logger("Removing callback with parameters: {} (needs validation data)".format(descriptor))
        else:
            kept.append(descriptor)
    callbacks[:] = kept
