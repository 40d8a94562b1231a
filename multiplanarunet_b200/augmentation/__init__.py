from .augmenters import Elastic2D  # noqa: F401
from .elastic_deformation import elastic_transform_2d  # noqa: F401
