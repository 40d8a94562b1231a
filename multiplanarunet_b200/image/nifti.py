"""Minimal NIfTI-1 (.nii / .nii.gz) reader and writer in numpy (nibabel is not a dependency of this
path).  Covers what mpunet's ImagePair needs from nibabel (mpunet/image/image_pair.py:163-192 reads
`get_fdata`-like arrays + `affine`; bin/predict.py:90-117 writes `<id>_PRED.nii.gz`): single-file
NIfTI-1, any scalar dtype, sform/qform affine, scl_slope/scl_inter scaling."""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8,
           512: np.uint16, 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).name: k for k, v in _DTYPES.items()}


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def _quaternion_affine(b, c, d, qfac, pixdim, offset):
    a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    z = pixdim.copy()
    z[2] *= -1.0 if qfac < 0 else 1.0
    A = np.eye(4)
    A[:3, :3] = R * z
    A[:3, 3] = offset
    return A


def read_nifti(path):
    """Returns (data ndarray in file dtype with scaling applied, affine 4x4 float64, header dict)."""
    with _open(path, "rb") as f:
        raw = f.read()
    for end in ("<", ">"):
        if struct.unpack(end + "i", raw[:4])[0] == 348:
            break
    else:
        raise ValueError("%s: not a NIfTI-1 file" % path)
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(end + "hh", raw[70:74])
    pixdim = np.array(struct.unpack(end + "8f", raw[76:108]), dtype=np.float64)
    vox_offset = struct.unpack(end + "f", raw[108:112])[0]
    slope, inter = struct.unpack(end + "ff", raw[112:120])
    qform_code, sform_code = struct.unpack(end + "hh", raw[252:256])
    quat = struct.unpack(end + "6f", raw[256:280])
    srow = np.array(struct.unpack(end + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    if raw[344:348] not in (b"n+1\x00", b"ni1\x00"):
        raise ValueError("%s: unsupported NIfTI magic %r" % (path, raw[344:348]))
    if datatype not in _DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype %d" % (path, datatype))
    ndim = dim[0]
    shape = tuple(int(d) for d in dim[1:1 + ndim])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(end)
    n = int(np.prod(shape))
    off = int(vox_offset) if vox_offset >= 352 else 352
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off).reshape(shape, order="F")
    data = np.ascontiguousarray(data.astype(dt.newbyteorder("=")))
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0 and np.isfinite(slope):
            data = data.astype(np.float64) * slope + inter
    if sform_code > 0:
        affine = np.eye(4)
        affine[:3, :] = srow
    elif qform_code > 0:
        affine = _quaternion_affine(quat[0], quat[1], quat[2], pixdim[0], pixdim[1:4].copy(), quat[3:6])
    else:
        affine = np.diag(list(pixdim[1:4]) + [1.0])
    return data, affine, {"dim": dim, "pixdim": pixdim, "datatype": datatype, "endianness": end}


def write_nifti(path, data, affine=None):
    data = np.asarray(data)
    if data.dtype.name not in _CODES:
        data = data.astype(np.float32)
    affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<hh", hdr, 70, _CODES[data.dtype.name], data.dtype.itemsize * 8)
    pix = np.linalg.norm(affine[:3, :3], axis=0)
    struct.pack_into("<8f", hdr, 76, 1.0, pix[0], pix[1], pix[2], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<ff", hdr, 112, 1.0, 0.0)
    hdr[123] = 2  # xyzt_units: mm
    struct.pack_into("<hh", hdr, 252, 0, 2)  # qform_code 0, sform_code 2 (aligned)
    struct.pack_into("<12f", hdr, 280, *affine[:3, :].ravel())
    hdr[344:348] = b"n+1\x00"
    with _open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(b"\x00" * 4)
        f.write(np.asfortranarray(data).tobytes(order="F"))
