"""Model registry, mirroring mpunet/models/__init__.py: classes are looked up by name from
hparams["build"]["model_class_name"] (mpunet/models/model_init.py:10-13)."""
from .unet import UNet  # noqa: F401
from .fusion_model import FusionModel  # noqa: F401
