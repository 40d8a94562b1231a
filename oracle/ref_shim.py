"""ORACLE tooling: import the UNMODIFIED reference (`/root/reference/mpunet`) in this container.

The reference's sampler / interpolator / mapping are plain numpy and run here under three stubs
(nothing under /root/reference is modified or copied):
  1. scipy.interpolate.interpnd._ndim_coords_from_arrays moved in scipy>=1.14
     (mpunet/interpolation/regular_grid_interpolator.py:3);
  2. nibabel is absent (mpunet/interpolation/sample_grid.py:1 only needs the import);
  3. tensorflow is absent (mpunet/sequences/base_sequence.py:2 needs keras.utils.Sequence).
Only oracle/make_golden.py and tests that pin the oracle use this; `/root/reference` does not exist on
the GPU box, so nothing that runs there may import this module.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MPUNET_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "mpunet"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    if not available():
        raise ImportError("reference tree not found at %s" % REF_ROOT)
    import scipy.interpolate
    if "scipy.interpolate.interpnd" not in sys.modules or not hasattr(
            sys.modules["scipy.interpolate.interpnd"], "_ndim_coords_from_arrays"):
        from scipy.interpolate import _interpnd
        _stub("scipy.interpolate.interpnd", _ndim_coords_from_arrays=_interpnd._ndim_coords_from_arrays)
        scipy.interpolate.interpnd = sys.modules["scipy.interpolate.interpnd"]
    if "nibabel" not in sys.modules:
        class _Hdr:  # pragma: no cover - attribute holder only
            quaternion_threshold = 0
        _stub("nibabel", Nifti1Header=_Hdr, load=lambda *a, **k: None)
    if "tensorflow" not in sys.modules:
        class _Sequence(object):
            pass
        tf = _stub("tensorflow")
        keras = _stub("tensorflow.keras")
        utils = _stub("tensorflow.keras.utils", Sequence=_Sequence)
        tf.keras = keras
        keras.utils = utils
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def modules():
    """Returns the reference modules the oracle is pinned against."""
    install()
    import importlib
    sg = importlib.import_module("mpunet.interpolation.sample_grid")
    vi = importlib.import_module("mpunet.interpolation.view_interpolator")
    rgi = importlib.import_module("mpunet.interpolation.regular_grid_interpolator")
    # fuse_and_predict imports mpunet.sequences.utils -> sequences package (needs the keras stub) and
    # mpunet.preprocessing (sklearn).  Import the module file directly to dodge mpunet.utils.fusion's
    # package __init__, which pulls mpunet.evaluate -> real TensorFlow.
    import importlib.util
    path = os.path.join(REF_ROOT, "mpunet", "utils", "fusion", "fuse_and_predict.py")
    # stub packages so that relative package imports resolve without executing their __init__
    spec = importlib.util.spec_from_file_location("_ref_fuse_and_predict", path)
    fap = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(fap)
    except Exception:  # pragma: no cover - depends on which optional deps are importable
        fap = None
    return types.SimpleNamespace(sample_grid=sg, view_interpolator=vi, rgi=rgi, fuse_and_predict=fap)
