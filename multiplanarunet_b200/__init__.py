"""multiplanarunet_b200 - B200-native implementation of the mpunet (perslev/MultiPlanarUNet) hot path.

Host code is Python mirroring the reference's plug points; all arithmetic runs in hand-written
sm_100a CUDA behind the C ABI declared in include/mpunet_b200.h (loaded by ``_C``).
"""
__version__ = "0.1.0"
