"""mpunet/errors/image_errors.py: both derive from AttributeError, as in the reference, so `getattr(im, "labels", None)`
style probing keeps working."""


class NoLabelFileError(AttributeError):
    pass


class ReadOnlyAttributeError(AttributeError):
    pass
