"""FusionModel on the GPU kernels, mirroring mpunet/models/fusion_model.py:14-75.

softmax(sum_v W[v,c] * x[n,v,c] + b[c]) with W init 1, b init 0 (fusion_model.py:21-39), trained with
the sparse generalized dice loss (uniform weights) + 1e-6*mean(w^2) regularisers (fusion_model.py:9-11,
55-56) and Adam (bin/train_fusion.py:345).  Gradient sums are all-reduced across ranks when
torch.distributed is initialised (points shard naturally).
"""
import ctypes

import numpy as np

from .. import _C
from .._C import lib, check


class FusionModel(object):
    def __init__(self, n_inputs, n_classes, weight="Simple", logger=None, verbose=True, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("FusionModel needs a CUDA device (no CPU fallback)")
        self.n_inputs, self.n_classes = int(n_inputs), int(n_classes)
        self.loss = "SparseGeneralizedDiceLoss(%s)" % weight
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.W = torch.ones(self.n_inputs, self.n_classes, dtype=torch.float32, device=self.device)
        self.b = torch.zeros(self.n_classes, dtype=torch.float32, device=self.device)
        self._m = torch.zeros(self.n_inputs * self.n_classes + self.n_classes, dtype=torch.float32, device=self.device)
        self._v = torch.zeros_like(self._m)
        self._accum = torch.zeros(self.n_inputs * self.n_classes + self.n_classes + 1, dtype=torch.float64,
                                  device=self.device)
        self.lr, self.beta_1, self.beta_2, self.epsilon, self.reg = 1e-3, 0.9, 0.999, 1e-7, 1e-6
        self.iterations = 0
        self.stop_training = False

    def compile(self, optimizer=None, loss=None, metrics=None, **kw):
        if optimizer is not None and hasattr(optimizer, "lr"):
            self.lr = float(optimizer.lr)
        return self

    def get_weights(self):
        return [self.W.cpu().numpy().copy(), self.b.cpu().numpy().reshape(1, -1).copy()]

    def set_weights(self, ws):
        import torch
        self.W.copy_(torch.as_tensor(np.asarray(ws[0], dtype=np.float32)))
        self.b.copy_(torch.as_tensor(np.asarray(ws[1], dtype=np.float32).reshape(-1)))

    def save_weights(self, path, overwrite=True):
        W, b = self.get_weights()
        if path.endswith((".h5", ".hdf5")):  # Keras layout of the reference's FusionLayer (fusion_model.py:14-43)
            from ..utils.keras_h5 import save_keras_weights
            save_keras_weights(path, {"fusion_layer": {"W": W, "b": b}})
            return
        with open(path, "wb") as f:
            np.savez(f, W=W, b=b)

    def load_weights(self, path, by_name=True):
        if path.endswith((".h5", ".hdf5")):
            from ..utils.keras_h5 import load_keras_weights
            layers = [d for d in load_keras_weights(path).values() if "W" in d and "b" in d]
            if len(layers) != 1:
                raise ValueError("%s does not hold exactly one fusion layer" % path)
            self.set_weights([layers[0]["W"], layers[0]["b"]])
            return
        with np.load(path) as z:
            self.set_weights([z["W"], z["b"]])

    def predict(self, x, batch_size=10000, verbose=0):
        """x [N,V,C] float32 -> softmax probabilities [N,C] (FusionLayer.call)."""
        raise NotImplementedError("fusion inference is fused with the mapping gather: use "
                                  "multiplanarunet_b200.utils.fusion.predict_multi_view")

    def train_on_batch(self, X, y, all_reduce=True):
        """One Adam step on a batch of points; X [n,V,C] f32 tensor (device), y [n] uint8 tensor.
        Returns the mean loss (incl. regulariser) over the GLOBAL batch."""
        import torch
        n = X.shape[0]
        self._accum.zero_()
        check(lib.mpu_fusion_grad(_C.ptr(X), _C.ptr(y), ctypes.c_longlong(n), self.n_inputs, self.n_classes,
                                  _C.ptr(self.W), _C.ptr(self.b), _C.ptr(self._accum), _C.current_stream()),
              "mpu_fusion_grad")
        n_total = float(n)
        if all_reduce and torch.distributed.is_available() and torch.distributed.is_initialized():
            cnt = torch.tensor([float(n)], dtype=torch.float64, device=self.device)
            torch.distributed.all_reduce(self._accum)
            torch.distributed.all_reduce(cnt)
            n_total = float(cnt.item())
        self.iterations += 1
        check(lib.mpu_fusion_adam(_C.ptr(self.W), _C.ptr(self.b), _C.ptr(self._m), _C.ptr(self._v),
                                  _C.ptr(self._accum), ctypes.c_double(n_total), self.n_inputs,
                                  self.n_classes, ctypes.c_float(self.reg), ctypes.c_float(self.lr),
                                  ctypes.c_float(self.beta_1), ctypes.c_float(self.beta_2),
                                  ctypes.c_float(self.epsilon), int(self.iterations), _C.current_stream()),
              "mpu_fusion_adam")
        return self._accum[-1] / n_total

    def fit(self, X, y, batch_size=2 ** 17, epochs=30, verbose=1, shuffle=True, **kw):
        """Epoch loop over device-resident points (bin/train_fusion.py:196-213)."""
        import torch
        n = X.shape[0]
        hist = []
        for ep in range(epochs):
            perm = torch.randperm(n, device=X.device) if shuffle else torch.arange(n, device=X.device)
            losses = []
            for s in range(0, n, batch_size):
                idx = perm[s:s + batch_size]
                losses.append(self.train_on_batch(X[idx].contiguous(), y[idx].contiguous()))
            hist.append(float(torch.stack(losses).mean().item()))
            if self.stop_training:
                break
        return hist
