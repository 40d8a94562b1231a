"""Plane-sampling front-end, mirroring mpunet/sequences/isotrophic_live_view_sequence_2d.py on the GPU
sampler: `get_view_from` builds whole inference stacks (reference :29-101, one launch instead of
276 numpy planes x 7 threads) and `__getitem__` draws a training batch of random oblique planes
(reference :119-216) with the foreground rules of isotrophic_live_view_sequence.py:70-128.

Image objects are duck-typed like the reference's ImagePair (image/image_pair.py): `.image`
[X,Y,Z,C] float32, `.labels` [X,Y,Z] uint8 | None, `.affine` 4x4, `.interpolator`
(multiplanarunet_b200.interpolation.ViewInterpolator), `.scaler_center/.scaler_scale` (RobustScaler
statistics, preprocessing/scaling.py:47-88), `.sample_weight`, `.predict_mode`.
"""
import numpy as np

from ..errors import ReadOnlyAttributeError
from ..interpolation import ViewInterpolator, plane_basis, plane_basis_batch, view_offsets


def robust_scaler_stats(image):
    """Per-channel (center_, scale_) of sklearn's RobustScaler fitted on ALL voxels of a [X,Y,Z,C] float32 volume, as
    MultiChannelScaler.fit does (preprocessing/scaling.py:47-73; sklearn.preprocessing.RobustScaler.fit: nanmedian,
    nanpercentile(25, 75), zero ranges -> 1).  float64 [C] each; pinned against the reference's fitted scaler in
    tests/golden/view_stack_*.npz."""
    center, scale = [], []
    for c in range(image.shape[-1]):
        col = np.ascontiguousarray(image[..., c], dtype=np.float32).reshape(-1, 1)
        center.append(float(np.nanmedian(col, axis=0)[0]))
        q = np.nanpercentile(col, (25.0, 75.0), axis=0)
        sc = float(q[1][0] - q[0][0])
        scale.append(sc if sc != 0.0 and abs(sc) >= 10 * np.finfo(np.float64).eps else 1.0)
    return np.asarray(center, dtype=np.float64), np.asarray(scale, dtype=np.float64)


class SyntheticImage(object):
    """Minimal ImagePair stand-in (what the sequence needs from image/image_pair.py:27-484):
    bg_value = 1st percentile per channel (bin/defaults/MultiPlanar/train_hparams.yaml:130,
    image_pair.py:469-484), scaler = RobustScaler median / IQR over all voxels."""

    def __init__(self, image, labels=None, affine=None, bg_value="1pct", bg_class=0, sample_weight=1.0,
                 device=None, identifier="synthetic"):
        image = np.asarray(image, dtype=np.float32)
        if image.ndim == 3:
            image = image[..., None]
        self._image = image
        self._labels = None if labels is None else np.asarray(labels).astype(np.uint8)
        self._affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
        self.shape = image.shape
        self.n_channels = image.shape[-1]
        self.predict_mode = labels is None
        self.sample_weight = float(sample_weight)
        self.identifier = identifier
        if isinstance(bg_value, str) and bg_value.endswith("pct"):
            pct = float(bg_value[:-3])
            bg_value = [float(np.percentile(image[..., c], pct)) for c in range(self.n_channels)]
        self.bg_value = bg_value
        self.bg_class = bg_class
        self.scaler_center, self.scaler_scale = robust_scaler_stats(image)
        # the device-resident interpolator is built on first use (needs a CUDA device; geometry and file handling of an
        # ImagePair do not)
        self._device = device
        self._interpolator = None

    # image / labels / affine are read-only like the reference's (image_pair.py:143-198)
    @property
    def image(self):
        return self._image

    @image.setter
    def image(self, _):
        raise ReadOnlyAttributeError("Manually setting the image attribute is not allowed. "
                                     "Initialize a new ImagePair object.")

    @property
    def labels(self):
        return self._labels

    @labels.setter
    def labels(self, _):
        raise ReadOnlyAttributeError("Manually setting the labels attribute is not allowed. "
                                     "Initialize a new ImagePair object.")

    @property
    def affine(self):
        return self._affine

    @affine.setter
    def affine(self, _):
        raise ReadOnlyAttributeError("Manually setting the affine attribute is not allowed. "
                                     "Initialize a new ImagePair object.")

    @property
    def interpolator(self):
        if self._interpolator is None:
            self._interpolator = ViewInterpolator(self._image, self._labels, self._affine, bg_value=self.bg_value,
                                                  bg_class=self.bg_class, device=self._device)
        return self._interpolator

    @interpolator.setter
    def interpolator(self, value):
        self._interpolator = value

    @property
    def center(self):
        """Voxel-space centre (image_pair.py:234-240)."""
        return (np.asarray(self.shape[:-1]) - 1) / 2

    @property
    def real_center(self):
        """Scanner-space centre (image_pair.py:242-248)."""
        return self._affine[:3, :3].dot(self.center) + self._affine[:3, -1]

    @property
    def real_shape(self):
        """Physical extent per axis: voxel counts x voxel sizes (image_pair.py:261-267, sample_grid.py:9-16; the voxel
        sizes are the column norms of the affine, which is what a NIfTI header's pixdim holds)."""
        pix = np.linalg.norm(self._affine[:3, :3], axis=0)
        return np.asarray(self.shape[:3]) * pix


class IsotrophicLiveViewSequence2D(object):
    def __init__(self, images, views, sample_dim, real_space_span, n_classes, batch_size=16,
                 noise_sd=0.1, fg_batch_fraction=0.5, force_all_fg="auto", is_validation=False,
                 label_crop=None, logger=None, list_of_augmenters=None, **kwargs):
        self.images = list(images) if isinstance(images, (list, tuple)) else [images]
        self.views = np.asarray(views, dtype=np.float64)
        self.sample_dim = int(sample_dim)
        self.real_space_span = real_space_span
        self.n_classes = int(n_classes)
        self.batch_size = int(batch_size)
        self.noise_sd = 0.0 if is_validation else noise_sd
        self.fg_batch_fraction = fg_batch_fraction
        self.force_all_fg_switch = force_all_fg
        self.is_validation = is_validation
        self.fg_classes = np.arange(1, self.n_classes)
        self.label_crop = np.zeros((2, 2), dtype=int) if label_crop is None else label_crop
        self.logger = logger or (lambda *a, **k: None)
        # on-the-fly augmenters (sequences/utils.py:38-47); never applied to validation batches
        self.list_of_augmenters = None if is_validation else (list(list_of_augmenters or []) or None)

    def augment(self, batch_x, batch_y, batch_w, bg_values):
        """isotrophic_live_view_sequence.py:130-139: every augmenter sees the scaled batch and the UNSCALED
        per-image background values (reference behaviour, …_2d.py:195-208)."""
        if self.list_of_augmenters:
            for aug in self.list_of_augmenters:
                batch_x, batch_y, batch_w = aug(batch_x=batch_x, batch_y=batch_y, batch_w=batch_w,
                                                bg_values=bg_values)
        return batch_x, batch_y, batch_w

    # -- reference properties (isotrophic_live_view_sequence.py:70-90)
    @property
    def n_fg_slices(self):
        return int(np.ceil(self.batch_size * self.fg_batch_fraction))

    @property
    def force_all_fg(self):
        if isinstance(self.force_all_fg_switch, str) and self.force_all_fg_switch.lower() == "auto":
            return self.batch_size > len(self.fg_classes)
        return self.force_all_fg_switch

    # -- inference stacks --------------------------------------------------------------------------
    def get_view_stack_device(self, image, view, n_planes="same+20", out_padded=None, cpad=0,
                              want_f32=True, want_labels=True):
        """Device-native get_view_from: returns (X [n,dim,dim,C] f32 tensor|None, y tensor|None,
        (axis, axis, offsets), inv_basis, basis).  With `out_padded` the planes go straight into the
        U-Net's bf16 input tensor."""
        dim, span = self.sample_dim, self.real_space_span
        bounding = float(np.linalg.norm(np.asarray(image.real_shape) / 2)) if n_planes == "by_radius" else None
        offsets = view_offsets(dim, span, n_planes, bounding)
        basis = plane_basis(view, 0.)
        X, y = image.interpolator.sample_planes(
            basis, offsets, dim, span, center=image.scaler_center, scale=image.scaler_scale,
            out_padded=out_padded, cpad=cpad, want_f32=want_f32,
            want_labels=want_labels and not image.predict_mode)
        hd = span // 2
        g = np.linspace(-hd, hd, dim)
        return X, y, (g, g, offsets), np.linalg.inv(basis), basis

    def get_view_from(self, image, view, n_planes):
        """Reference layout (…_2d.py:29-101): X [dim,dim,n,C] float32, y [dim,dim,n] uint8 | None,
        (real_axis, real_axis, offsets), inv_basis."""
        X, y, grid, inv_basis, _ = self.get_view_stack_device(image, view, n_planes)
        Xn = np.ascontiguousarray(X.permute(1, 2, 0, 3).cpu().numpy())
        yn = None if y is None else np.ascontiguousarray(y.permute(1, 2, 0).cpu().numpy())
        return Xn, yn, grid, inv_basis

    # -- training batches --------------------------------------------------------------------------
    # -- the reference's acceptance rules on class-presence vectors (isotrophic_live_view_sequence.py:98-128) -------
    def _validate_lab_vec(self, present, has_fg, cur_bs):
        new_mask = has_fg + present
        if np.all(new_mask):
            return True, new_mask
        if np.sum(new_mask == 0) < (self.batch_size - cur_bs):
            return True, new_mask
        return False, has_fg

    def _validate_lab(self, present, has_fg_count, cur_bs):
        if np.any(present):
            return True, 1
        if (self.n_fg_slices - has_fg_count) < (self.batch_size - cur_bs):
            return True, 0
        return False, 0

    def draw_candidates(self, max_tries=10, rng=np.random):
        """Candidate planes of one batch, drawn like `_get_valid_slice_from` draws them (…_2d.py:128-144: view index,
        offset ~ U(-span//2, span//2), normal noise sd `noise_sd`) but for all B x max_tries tries up front.
        Returns (image index [B], view index [B,T], offsets [B,T], noise [B,T,3])."""
        B, T = self.batch_size, max_tries
        sphere_r = self.real_space_span // 2
        im_idx = rng.randint(0, len(self.images), B)
        view_idx = rng.randint(0, len(self.views), size=(B, T))
        offs = rng.uniform(-sphere_r, sphere_r, size=(B, T))
        noise = rng.normal(scale=self.noise_sd, size=(B, T, 3)) if self.noise_sd else np.zeros((B, T, 3))
        return im_idx, view_idx, offs, noise

    def select_slices(self, present, valid_im):
        """The sequential accept / reject loop of `_get_valid_slice_from` (…_2d.py:119-161) for every slot of a batch,
        on per-candidate facts: present [B,T,n_fg] bool = np.isin(fg_classes, lab), valid_im [B,T] bool = is_valid_im.
        As in the reference, the class-coverage vector `has_fg_vec` starts from zeros for EVERY slot (the caller's
        vector is never updated: `_get_valid_slice_from` rebinds a local) and accumulates over the tries of one slot;
        a candidate rejected by is_valid_im does not count its foreground.  Returns (picks [B], has_fg_count)."""
        B, T = present.shape[:2]
        picks = np.empty(B, dtype=np.int64)
        has_fg_count = 0
        for s in range(B):
            has_fg_vec = np.zeros_like(self.fg_classes)
            picks[s] = T - 1
            for t in range(T):
                last = t == T - 1
                if self.force_all_fg and not last:
                    ok, has_fg_vec = self._validate_lab_vec(present[s, t], has_fg_vec, s)
                    if not ok:
                        continue
                ok, change = self._validate_lab(present[s, t], has_fg_count, s)
                if ok or last:
                    if last or valid_im[s, t]:
                        has_fg_count += change
                        picks[s] = t
                        break
        return picks, has_fg_count

    def sample_batch_device(self, out_padded=None, cpad=0, max_tries=10, rng=np.random, candidates=None,
                            return_picks=False):
        """One training batch of random oblique planes (the reference's __getitem__, …_2d.py:163-216).  All
        B x max_tries candidate planes are probed in ONE launch per image (mpu_probe_planes: class-presence bits from
        the nearest-label gather, `is_valid_im` from the trilinear image) and the two small flag arrays come back in a
        single copy; the reference's sequential accept / reject rules run on them on the host, and only the accepted
        planes are materialised (trilinear image + labels + RobustScaler), optionally straight into the U-Net's bf16
        input tensor.  `candidates` = (image index, view index, offsets, noise) replaces the random draws (tests).
        Returns (x [B,dim,dim,C] f32 tensor | None if out_padded, y [B,dim,dim] uint8 tensor, w [B])."""
        import torch
        B, dim, span = self.batch_size, self.sample_dim, self.real_space_span
        im_idx, view_idx, cand_offs, noise = candidates if candidates is not None else self.draw_candidates(max_tries, rng)
        T = view_idx.shape[1]
        cand_bases = plane_basis_batch(self.views[view_idx.ravel()], noise.reshape(-1, 3)).reshape(B, T, 3, 3)
        nfg = len(self.fg_classes)
        present = np.zeros((B, T, nfg), dtype=bool)
        valid_im = np.zeros((B, T), dtype=bool)
        pending = []
        for i, image in enumerate(self.images):
            slots = np.where(im_idx == i)[0]
            if len(slots) == 0:
                continue
            cm, vd = image.interpolator.probe_planes(cand_bases[slots].reshape(-1, 3, 3), cand_offs[slots].ravel(),
                                                     dim, span)
            pending.append((slots, torch.stack([cm, vd]).to("cpu", non_blocking=False)))
        for slots, flags in pending:
            flags = flags.numpy().astype(np.int64) & 0xFFFFFFFF
            cm = flags[0].reshape(len(slots), T)
            present[slots] = ((cm[..., None] >> np.asarray(self.fg_classes)[None, None, :]) & 1).astype(bool)
            valid_im[slots] = flags[1].reshape(len(slots), T) != 0
        picks, _ = self.select_slices(present, valid_im)
        ar = np.arange(B)
        chosen_bases, chosen_offs = cand_bases[ar, picks], cand_offs[ar, picks]
        if self.list_of_augmenters and out_padded is not None:
            raise NotImplementedError("augmenters need the float32 batch: call sample_batch_device() without "
                                      "out_padded and pack the augmented batch")
        dev = self.images[0].interpolator.device
        x = torch.empty(B, dim, dim, self.images[0].n_channels, dtype=torch.float32,
                        device=dev) if out_padded is None else None
        y = torch.empty(B, dim, dim, dtype=torch.uint8, device=dev)
        w = np.empty(B, dtype=np.float32)
        for i, image in enumerate(self.images):
            slots = np.where(im_idx == i)[0]
            if len(slots) == 0:
                continue
            contiguous = len(self.images) == 1
            xi, yi = image.interpolator.sample_planes(
                chosen_bases[slots], chosen_offs[slots], dim, span, center=image.scaler_center,
                scale=image.scaler_scale, out_padded=out_padded if contiguous else None, cpad=cpad,
                want_f32=(out_padded is None) or not contiguous, want_labels=True,
                out_f32=x if (contiguous and x is not None) else None, out_labels=y if contiguous else None)
            if not contiguous:
                if x is None:
                    raise NotImplementedError("direct U-Net input writes need a single resident image per rank")
                x[torch.as_tensor(slots, device=dev)] = xi
                y[torch.as_tensor(slots, device=dev)] = yi
            w[slots] = image.sample_weight
        if self.list_of_augmenters:
            bg_values = [self.images[i].interpolator.bg_value for i in im_idx]
            wl = list(w)
            x, y, wl = self.augment(x, y, wl, bg_values)
            w = np.asarray(wl, dtype=np.float32)
        if return_picks:
            return x, y, w, picks
        return x, y, w

    def prefetched(self, depth=2, max_tries=10, rng=None):
        """Endless iterator of training batches (x, y, w) sampled AHEAD of the consumer: a worker thread runs
        sample_batch_device on its own CUDA stream (candidate probe, flag read-back, accepted planes, augmenters), so
        the rejection sampler's host round trip overlaps the previous train step instead of serialising with it
        (the reference hides its CPU sampler behind 5 Keras generator workers, train/trainer.py:246-257)."""
        return BatchPrefetcher(self, depth, max_tries, rng)

    def __getitem__(self, idx):
        """Reference batch layout (…_2d.py:163-216 + prepare_batches): x [B,dim,dim,C] f32,
        y [B,dim*dim,1] uint8, w [B]."""
        x, y, w = self.sample_batch_device()
        B = x.shape[0]
        return x.cpu().numpy(), y.cpu().numpy().reshape(B, -1, 1), w.astype(np.float64)

    def __len__(self):
        return 10 ** 6


class BatchPrefetcher(object):
    """Background producer of training batches; see IsotrophicLiveViewSequence2D.prefetched."""

    def __init__(self, seq, depth=2, max_tries=10, rng=None):
        import queue
        import threading
        import torch
        self.seq, self.max_tries = seq, max_tries
        self.rng = rng if rng is not None else np.random.RandomState(np.random.randint(0, 2 ** 31 - 1))
        self.device = seq.images[0].interpolator.device
        self.stream = torch.cuda.Stream(device=self.device)
        self.q = queue.Queue(maxsize=depth)
        self._stop = False
        self._err = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        import torch
        try:
            torch.cuda.set_device(self.device)
            while not self._stop:
                with torch.cuda.stream(self.stream):
                    x, y, w = self.seq.sample_batch_device(max_tries=self.max_tries, rng=self.rng)
                    wt = torch.as_tensor(np.asarray(w, dtype=np.float32)).to(self.device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.stream)
                while not self._stop:
                    try:
                        self.q.put((x, y, wt, ev), timeout=0.2)
                        break
                    except Exception:
                        continue
        except Exception as e:  # surfaced to the consumer
            self._err = e
            self.q.put(None)

    def __iter__(self):
        return self

    def __next__(self):
        import torch
        item = self.q.get()
        if item is None:
            raise self._err
        x, y, w, ev = item
        torch.cuda.current_stream(self.device).wait_event(ev)
        for t in (x, y, w):  # the tensors were allocated on the worker's stream
            t.record_stream(torch.cuda.current_stream(self.device))
        return x, y, w

    def close(self):
        self._stop = True
        try:
            while True:
                self.q.get_nowait()
        except Exception:
            pass
        self.thread.join(timeout=2)
