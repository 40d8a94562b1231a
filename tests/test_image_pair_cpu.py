"""ImagePair through the class, on the values the reference's only numeric test pins
(mpunet/tests/integration/test_image_pair_with_valid_image.py:12-108): a 12x14x16x3 float64 volume with affine
diag(1, 0.5, 0.1, 1) saved as .nii.gz.  nibabel is not installed here, so the file is laid out in this test directly from
the NIfTI-1 specification (348-byte header, field offsets as in nifti1.h) - NOT with the package's own writer - in the
variants nibabel itself produces (little-endian, sform + qform set) plus a big-endian, qform-only one with a non-empty
header extension."""
import gzip
import struct

import numpy as np
import pytest

from multiplanarunet_b200.errors import ReadOnlyAttributeError
from multiplanarunet_b200.image import ImagePair
from multiplanarunet_b200.image.nifti import read_nifti

DATA = np.random.RandomState(3).randn(12, 14, 16, 3).astype(np.float64)
AFFINE = np.diag([1, 0.5, 0.1, 1])


def nifti1_bytes(data, affine, end="<", sform=True, qform=True, extension=b""):
    """A single-file NIfTI-1 image following nifti1.h; float64 voxels (datatype 64), mm units."""
    hdr = bytearray(348)
    struct.pack_into(end + "i", hdr, 0, 348)                                   # sizeof_hdr
    struct.pack_into(end + "8h", hdr, 40, data.ndim, *data.shape, *([1] * (7 - data.ndim)))   # dim[8]
    struct.pack_into(end + "hh", hdr, 70, 64, 64)                              # datatype FLOAT64, bitpix
    pix = np.linalg.norm(affine[:3, :3], axis=0)
    struct.pack_into(end + "8f", hdr, 76, 1.0, *pix, 1.0, 1.0, 1.0, 1.0)       # pixdim[8], qfac = +1
    vox_offset = 352 + len(extension)
    struct.pack_into(end + "f", hdr, 108, float(vox_offset))
    struct.pack_into(end + "ff", hdr, 112, 1.0, 0.0)                           # scl_slope, scl_inter
    hdr[123] = 2                                                               # xyzt_units
    struct.pack_into(end + "hh", hdr, 252, 1 if qform else 0, 2 if sform else 0)
    struct.pack_into(end + "6f", hdr, 256, 0.0, 0.0, 0.0, *affine[:3, 3])      # quatern b c d (identity rotation), offsets
    if sform:
        struct.pack_into(end + "12f", hdr, 280, *affine[:3, :].ravel())
    hdr[344:348] = b"n+1\x00"
    ext_flag = (b"\x01" if extension else b"\x00") + b"\x00" * 3
    vox = np.asfortranarray(data.astype(np.dtype(np.float64).newbyteorder(end))).tobytes(order="F")
    return bytes(hdr) + ext_flag + extension + vox


def write(path, **kw):
    with gzip.open(path, "wb") as f:
        f.write(nifti1_bytes(DATA, AFFINE, **kw))
    return str(path)


VARIANTS = {
    "nibabel_like": dict(),
    "big_endian_qform_only_with_extension": dict(end=">", sform=False, qform=True,
                                                 extension=struct.pack(">ii", 16, 4) + b"comment\x00"),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_stored_image_matches_disk_image(tmp_path, variant):
    im = ImagePair(img_path=write(tmp_path / "valid_image.nii.gz", **VARIANTS[variant]))
    assert im.predict_mode                                    # initialised with only an image
    assert list(im.shape) == list(DATA.shape)
    assert np.isclose(im.image, DATA).all()
    assert im.image.dtype == np.dtype("float32")              # ImagePair always stores float32 images
    assert np.isclose(im.affine, AFFINE).all()
    assert im.labels is None                                  # image_pair.py:185-192 (the reference's CODE; its stale
                                                              # test expects NoLabelFileError, which :190 swallows)
    assert im.identifier == "valid_image"


def test_error_raising(tmp_path):
    im = ImagePair(img_path=write(tmp_path / "valid_image.nii.gz"))
    for attr in ("image", "labels", "affine"):
        with pytest.raises(ReadOnlyAttributeError):
            setattr(im, attr, [1, 2, 3])


def test_shape_values(tmp_path):
    im = ImagePair(img_path=write(tmp_path / "valid_image.nii.gz"))
    assert list(im.center) == [5.5, 6.5, 7.5]                               # voxel-space centre (zero-indexed)
    assert np.isclose(np.array(im.real_center), [5.5, 3.25, 0.75]).all()    # scanner-space centre
    assert np.isclose(np.array(im.real_shape), [12, 7, 1.6]).all()          # mm in scanner space


def test_reader_scaling_and_label_pair(tmp_path):
    """scl_slope / scl_inter are applied like nibabel's get_fdata; labels load as uint8 next to the image."""
    raw = bytearray(nifti1_bytes(DATA, AFFINE))
    struct.pack_into("<ff", raw, 112, 2.0, -1.0)
    p = tmp_path / "scaled.nii"
    p.write_bytes(bytes(raw))
    got, aff, _ = read_nifti(str(p))
    assert np.allclose(got, DATA * 2.0 - 1.0) and np.allclose(aff, AFFINE)
    lab = (np.abs(DATA[..., 0]) * 2).astype(np.float64)       # label maps are often stored as floats
    pl = tmp_path / "labels.nii"
    pl.write_bytes(nifti1_bytes(lab, AFFINE))
    im = ImagePair(img_path=str(p), labels_path=str(pl))
    assert not im.predict_mode and im.labels.dtype == np.uint8 and im.labels.shape == DATA.shape[:3]
    assert np.array_equal(im.labels, lab.astype(np.uint8))
