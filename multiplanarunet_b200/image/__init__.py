from .nifti import read_nifti, write_nifti  # noqa: F401
from .image_pair import ImagePair, ImagePairLoader, Auditor  # noqa: F401
