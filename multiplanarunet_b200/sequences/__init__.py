from .isotrophic_live_view_sequence_2d import IsotrophicLiveViewSequence2D, SyntheticImage  # noqa: F401
from .utils import get_sequence  # noqa: F401
