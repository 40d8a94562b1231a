"""ctypes loader for libmpunet_b200.so (the C-ABI declared in include/mpunet_b200.h).

There is deliberately no fallback: if the shared library is missing or a symbol is absent the import
fails loudly, and compute entry points return an error without a CUDA device.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpunet_b200.so")


class MpuError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libmpunet_b200.so not found at %s - build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C multiplanarunet_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()
lib.mpu_last_error.restype = ctypes.c_char_p
lib.mpu_version.restype = ctypes.c_int


def check(rc, what=""):
    if rc != 0:
        msg = lib.mpu_last_error()
        raise MpuError("%s failed (%d): %s" % (what or "mpunet_b200 call", rc,
                                               msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def float_array(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def double_array(vals):
    return (ctypes.c_double * len(vals))(*[float(v) for v in vals])


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    """There is no CPU path: every compute entry point needs a CUDA device."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("multiplanarunet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
