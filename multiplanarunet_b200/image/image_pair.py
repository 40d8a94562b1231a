"""ImagePair / ImagePairLoader / Auditor equivalents on the device sampler (mirrors of
mpunet/image/image_pair.py:27-484, image_pair_loader.py:18-358, auditor.py:37-260 reduced to what the
MultiPlanar hot path consumes)."""
import glob
import os

import numpy as np

from .nifti import read_nifti
from ..sequences.isotrophic_live_view_sequence_2d import SyntheticImage


class ImagePair(SyntheticImage):
    """One image (+ optional label map) of a project: loaded from NIfTI, kept resident on the GPU."""

    def __init__(self, img_path, labels_path=None, sample_weight=1.0, bg_class=0, bg_value="1pct",
                 device=None, logger=None):
        data, affine, _ = read_nifti(img_path)
        data = np.asarray(data, dtype=np.float32)
        if data.ndim == 3:
            data = data[..., None]
        labels = None
        if labels_path is not None:
            lab, _, _ = read_nifti(labels_path)
            labels = np.asarray(lab).squeeze().astype(np.uint8)
        base = os.path.basename(img_path)
        ident = base.split(".nii")[0]
        super().__init__(data, labels, affine, bg_value=bg_value, bg_class=bg_class,
                         sample_weight=sample_weight, device=device, identifier=ident)
        self.image_path, self.labels_path = img_path, labels_path


class ImagePairLoader(object):
    def __init__(self, base_dir="./", img_subdir="images", label_subdir="labels", sample_weight=1.0,
                 bg_class=0, predict_mode=False, bg_value="1pct", device=None, logger=None, **kwargs):
        self.base_dir = os.path.abspath(base_dir) if base_dir else None
        self.img_dir = os.path.join(self.base_dir, img_subdir) if self.base_dir else None
        self.lab_dir = os.path.join(self.base_dir, label_subdir) if self.base_dir and label_subdir else None
        self.predict_mode = predict_mode or not (self.lab_dir and os.path.isdir(self.lab_dir))
        self.kw = dict(sample_weight=sample_weight, bg_class=bg_class, bg_value=bg_value, device=device)
        self.image_paths = sorted(glob.glob(os.path.join(self.img_dir, "*.nii*"))) if self.img_dir else []
        self._cache = {}

    def __len__(self):
        return len(self.image_paths)

    def label_path_for(self, img_path):
        if self.predict_mode:
            return None
        p = os.path.join(self.lab_dir, os.path.basename(img_path))
        return p if os.path.exists(p) else None

    def get(self, i):
        if i not in self._cache:
            p = self.image_paths[i]
            self._cache[i] = ImagePair(p, self.label_path_for(p), **self.kw)
        return self._cache[i]

    @property
    def images(self):
        return [self.get(i) for i in range(len(self))]

    def unload(self, i):
        self._cache.pop(i, None)


class Auditor(object):
    """Fills dim / real_space_span / n_channels / n_classes from the data when the YAML has them Null
    (auditor.py:73-125,199-209 heuristics: span = 75th percentile of real sizes, resolution = 25th
    percentile of voxel sizes, dim = nearest multiple of 16 in [128,512])."""

    def __init__(self, nii_paths, nii_lab_paths=None, min_dim_2d=128, max_dim_2d=512, span_percentile=75,
                 res_percentile=25, hparams=None):
        shapes, sizes, pix, chans = [], [], [], []
        for p in nii_paths:
            data, affine, _ = read_nifti(p)
            pd = np.linalg.norm(affine[:3, :3], axis=0)
            shapes.append(data.shape[:3])
            sizes.append(np.asarray(data.shape[:3]) * pd)
            pix.append(pd)
            chans.append(data.shape[3] if data.ndim > 3 else 1)
        self.n_channels = int(chans[0])
        self.n_classes = hparams.get_from_anywhere("n_classes") if hparams is not None else None
        if self.n_classes is None and nii_lab_paths:
            mx = 0
            for p in nii_lab_paths:
                lab, _, _ = read_nifti(p)
                mx = max(mx, int(np.max(lab)))
            self.n_classes = mx + 1
        span = np.percentile(sizes, span_percentile)
        res = np.percentile(pix, res_percentile)
        valid = np.array([i for i in range(min_dim_2d, max_dim_2d + 1) if (i * 0.5 ** 4).is_integer()])
        sample_dim = span / res
        nearest = valid[np.abs(valid - sample_dim).argmin()]
        if nearest < sample_dim * 0.90:
            span = max(int(span * 0.70), nearest * res)
        self.sample_dim_2D, self.real_space_span_2D = int(nearest), float(span)

    def fill(self, hparams, model_type="2d"):
        for group, name, value in (("fit", "real_space_span", self.real_space_span_2D),
                                   ("build", "dim", self.sample_dim_2D),
                                   ("build", "n_channels", self.n_channels),
                                   ("build", "n_classes", self.n_classes)):
            hparams.set_value(subdir=group, name=name, value=value)
        hparams.save_current()
