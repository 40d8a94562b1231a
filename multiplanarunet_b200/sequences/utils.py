"""`get_sequence`: the sequence registry of mpunet/sequences/utils.py:5-79 for the paths this build covers.

`intrp_style: iso_live` (the MultiPlanar default, train_hparams.yaml `fit.intrp_style`) selects the device-backed
IsotrophicLiveViewSequence2D; the 3D box / patch styles belong to the 3D model families, which are out of scope
 and are refused by name.  Augmenter descriptors ({cls_name, kwargs}) are turned into objects for
training sequences only, exactly like the reference."""
from .isotrophic_live_view_sequence_2d import IsotrophicLiveViewSequence2D

_OUT_OF_SCOPE = ("iso_live_3d", "patches_3d", "sliding_patches_3d")


def _images_of(data_queue):
    """ImagePairLoader-like (len + get / images), queue-like (.dataset) or a plain list of images."""
    if hasattr(data_queue, "dataset"):
        data_queue = data_queue.dataset
    if isinstance(data_queue, (list, tuple)):
        return list(data_queue)
    if hasattr(data_queue, "get") and hasattr(data_queue, "__len__"):
        return [data_queue.get(i) for i in range(len(data_queue))]
    if hasattr(data_queue, "images"):
        return list(data_queue.images)
    raise TypeError("cannot take images from %r" % (data_queue,))


def get_sequence(data_queue, is_validation, logger=None, augmenters=None, **seq_kwargs):
    logger = logger or print
    aug_list = []
    if not is_validation and augmenters:
        logger("Using on-the-fly augmenters:")
        from .. import augmentation
        for aug in augmenters:
            cls = augmentation.__dict__.get(aug["cls_name"])
            if cls is None:
                raise NotImplementedError("augmenter %r is not available on the B200 path" % aug["cls_name"])
            obj = cls(**aug["kwargs"])
            aug_list.append(obj)
            logger(obj)
    style = str(seq_kwargs.pop("intrp_style", "iso_live")).lower()
    if style == "iso_live":
        return IsotrophicLiveViewSequence2D(_images_of(data_queue), is_validation=is_validation,
                                            list_of_augmenters=aug_list, logger=logger, **seq_kwargs)
    if style in _OUT_OF_SCOPE:
        raise NotImplementedError("intrp_style '%s' belongs to the 3D model families, which this build does not "
                                  "cover (2D multi-planar path only)" % style)
    raise ValueError("Invalid interpolator schema '%s' specified" % style)
