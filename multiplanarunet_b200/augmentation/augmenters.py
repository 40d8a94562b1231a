"""On-the-fly augmenters selected by name from train_hparams.yaml (`augmenters: [{cls_name, kwargs}]`,
mpunet/sequences/utils.py:38-47; classes mpunet/augmentation/augmenters.py:5-126).

Elastic2D runs on the device over the whole batch; the random draws follow the reference's order (mask for the
batch, then per augmented slice: alpha, sigma, dx noise, dy noise) with numpy's global generator."""
import numpy as np

from .elastic_deformation import elastic_batch_device


class Augmenter(object):
    def __call__(self, batch_x, batch_y, bg_values, batch_w=None):
        raise NotImplementedError


class Elastic(Augmenter):
    def __init__(self, alpha, sigma, apply_prob, aug_weight=0.33):
        super().__init__()
        if isinstance(alpha, (list, tuple)):
            if len(alpha) != 2:
                raise ValueError("Invalid list of alphas specified '%s'. Should be 2 numbers." % (alpha,))
            if alpha[1] <= alpha[0]:
                raise ValueError("alpha upper is smaller than sigma lower (%s)" % (alpha,))
        if isinstance(sigma, (list, tuple)):
            if len(sigma) != 2:
                raise ValueError("Invalid list of sigmas specified '%s'. Should be 2 numbers." % (sigma,))
            if sigma[1] <= sigma[0]:
                raise ValueError("Sigma upper is smaller than sigma lower (%s)" % (sigma,))
        if apply_prob > 1 or apply_prob < 0:
            raise ValueError("Apply probability is invalid with value %3.f" % apply_prob)
        self._alpha = alpha
        self._sigma = sigma
        self.apply_prob = apply_prob
        self.weight = aug_weight
        self.__name__ = "Elastic"

    @property
    def alpha(self):
        if isinstance(self._alpha, (list, tuple)):
            return np.random.uniform(self._alpha[0], self._alpha[1], 1)[0]
        return self._alpha

    @property
    def sigma(self):
        if isinstance(self._sigma, (list, tuple)):
            return np.random.uniform(self._sigma[0], self._sigma[1], 1)[0]
        return self._sigma

    def __str__(self):
        return "%s(alpha=%s, sigma=%s, apply_prob=%.3f)" % (self.__name__, self._alpha, self._sigma,
                                                            self.apply_prob)

    __repr__ = __str__


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
