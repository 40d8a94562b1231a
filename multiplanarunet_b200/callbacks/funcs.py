"""mpunet/callbacks/funcs.py:5-84: callback descriptors from the YAML -> objects."""
from .callbacks import DelayedCallback


def init_callback_objects(callbacks, logger):
    """callbacks: list of {"class_name", "kwargs"[, "start_from", "pass_logger"]} dicts or initialised objects.
    Returns (objects, {class_name: object})."""
    from . import callbacks as tcb
    from . import validation as tval
    cb_objs, cb_dict = [], {}
    for i, callback in enumerate(callbacks):
        if not isinstance(callback, dict):
            cb = callback
            kwargs = {"params": "?"}
            cls_name = callback.__class__.__name__
            start_from = 0
        else:
            kwargs = dict(callback.get("kwargs") or {})
            cls_name = callback["class_name"]
            start_from = callback.get("start_from")
            if callback.get("pass_logger"):
                kwargs["logger"] = logger
            cls = getattr(tcb, cls_name, None) or getattr(tval, cls_name, None)
            if cls is None:
                raise ValueError("No callback named %s" % cls_name)
            cb = cls(**kwargs)
        if start_from:
            logger("OBS: '%s' activates at epoch %i" % (cls_name, start_from))
            cb = DelayedCallback(callback=cb, start_from=start_from, logger=logger)
        cb_objs.append(cb)
        cb_dict[cls_name] = cb
        logger("[%i] Using callback: %s(%s)" % (i + 1, cb.__class__.__name__,
                                                ", ".join(["%s=%s" % (a, kwargs[a]) for a in kwargs])))
    return cb_objs, cb_dict


def remove_validation_callbacks(callbacks, logger=None):
    """Drops every descriptor with a 'val'-mentioning kwarg (needs validation data).  The reference pops while
    enumerating and can skip the entry that follows a removed one (funcs.py:72-84); this keeps the intent."""
    kept = []
    for callback in callbacks:
        needs_val = isinstance(callback, dict) and any(
            "val" in str(p).lower() for p in (callback.get("kwargs") or {}).values())
        if needs_val:
            if logger:
                logger("Removing callback with parameters: {} (needs validation data)".format(callback))
        else:
            kept.append(callback)
    callbacks[:] = kept
