"""Exception types of the image layer (mpunet/errors/image_errors.py)."""
from .image_errors import NoLabelFileError, ReadOnlyAttributeError  # noqa: F401
