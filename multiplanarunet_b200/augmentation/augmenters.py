"""On-the-fly augmenters selected by name from train_hparams.yaml (`augmenters: [{cls_name, kwargs}]`,
mpunet/sequences/utils.py:38-47; classes mpunet/augmentation/augmenters.py:5-126).

Elastic2D runs on the device over the whole batch; the random draws follow the reference's order (mask for the
batch, then per augmented slice: alpha, sigma, dx noise, dy noise) with numpy's global generator."""
import numpy as np

from .elastic_deformation import elastic_batch_device


class Augmenter(object):
    def __call__(self, batch_x, batch_y, bg_values, batch_w=None):
        raise NotImplementedError


def _checked_range(value, what):
    """A scalar stays a scalar; a pair must be (low, high) with low < high (the YAML's `[0, 450]` style)."""
    if not isinstance(value, (list, tuple)):
        return value
    if len(value) != 2:
        raise ValueError("%s must be a number or a [low, high] pair, got %r" % (what, value))
    low, high = value
    if not low < high:
        raise ValueError("%s range %r is empty (low must be below high)" % (what, value))
    return (low, high)


def _draw(value):
    """Scalar: itself; (low, high): one uniform draw from numpy's global generator (same call as the reference, so a
    seeded run consumes the stream identically: np.random.uniform(low, high, 1)[0])."""
    if isinstance(value, tuple):
        return np.random.uniform(value[0], value[1], 1)[0]
    return value


class Elastic(Augmenter):
    """Parameters of the elastic deformation augmenters (mpunet/augmentation/augmenters.py:41-84): displacement
    strength `alpha` and smoothness `sigma`, each fixed or drawn per slice from a range, the probability of
    deforming a slice, and the sample weight given to deformed slices."""

    def __init__(self, alpha, sigma, apply_prob, aug_weight=0.33):
        super().__init__()
        self._alpha = _checked_range(alpha, "alpha")
        self._sigma = _checked_range(sigma, "sigma")
        if not 0 <= apply_prob <= 1:
            raise ValueError("apply_prob must lie in [0, 1], got %r" % (apply_prob,))
        self.apply_prob = apply_prob
        self.weight = aug_weight
        self.__name__ = "Elastic"

    alpha = property(lambda self: _draw(self._alpha))
    sigma = property(lambda self: _draw(self._sigma))

    def __repr__(self):
        return "%s(alpha=%s, sigma=%s, apply_prob=%.3f)" % (self.__name__, list(self._alpha) if isinstance(
            self._alpha, tuple) else self._alpha, list(self._sigma) if isinstance(self._sigma, tuple) else self._sigma,
            self.apply_prob)

    __str__ = __repr__


class Elastic2D(Elastic):
    """Random elastic deformation of some slices of a batch (linear for images, nearest for labels).

    batch_x: device tensor [B,H,W,C] float32 (or a list of numpy [H,W,C] arrays, as in the reference),
    batch_y: device tensor [B,H,W] uint8 (or list of numpy), bg_values: per image list of per-channel fills,
    batch_w: optional per-slice weights (augmented slices get `aug_weight`)."""

    def __init__(self, alpha, sigma, apply_prob, aug_weight=0.33):
        super().__init__(alpha, sigma, apply_prob, aug_weight)
        self.__name__ = "Elastic2D"

    def __call__(self, batch_x, batch_y, bg_values, batch_w=None):
        import torch
        as_lists = not torch.is_tensor(batch_x)
        if as_lists:
            xs = torch.as_tensor(np.stack([np.asarray(x, dtype=np.float32).reshape(
                np.asarray(x).shape[:2] + (-1,)) for x in batch_x])).cuda()
            ys = torch.as_tensor(np.stack([np.asarray(y).astype(np.uint8) for y in batch_y])).cuda()
        else:
            xs, ys = batch_x, batch_y
        B, H, W, C = xs.shape
        mask = np.random.rand(B) <= self.apply_prob
        idx = np.nonzero(mask)[0]
        if len(idx):
            alphas, sigmas, noise = [], [], np.empty((len(idx), 2, H, W), dtype=np.float64)
            for k in range(len(idx)):
                alphas.append(self.alpha)
                sigmas.append(self.sigma)
                noise[k, 0] = np.random.rand(H, W) * 2 - 1
                noise[k, 1] = np.random.rand(H, W) * 2 - 1
            sel = torch.as_tensor(idx, device=xs.device)
            bgs = [list(np.broadcast_to(np.asarray(bg_values[i], dtype=np.float32).ravel(), (C,))) for i in idx]
            xo, yo = elastic_batch_device(xs[sel], ys[sel].reshape(len(idx), H, W) if ys is not None else None,
                                          noise, alphas, sigmas, bgs)
            xs = xs.clone()
            xs[sel] = xo
            if ys is not None:
                shp = ys.shape
                ys = ys.clone().reshape(B, H, W)
                ys[sel] = yo
                ys = ys.reshape(shp)
            if batch_w is not None:
                for i in idx:
                    batch_w[i] = self.weight
        if as_lists:
            xs = [x for x in xs.cpu().numpy()]
            ys = [y.astype(np.asarray(batch_y[0]).dtype) for y in ys.cpu().numpy()]
        if batch_w is not None:
            return xs, ys, batch_w
        return xs, ys
