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

    def _distributed(self):
        import torch
        return torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size() > 1

    def enable_peer_exchange(self, group=None):
        """Multi-rank training with the gradient exchange fused into the train-step kernel: allocates this rank's
        mailbox in peer-mapped memory (torch.distributed._symmetric_memory: CUDA IPC / fabric handles over NVLink) and
        exchanges the pointers.  Afterwards fit() runs mpu_fusion_train_epoch_peer - one launch per batch, no NCCL
        call per step.  Returns False (and keeps the NCCL path) when symmetric memory is unavailable."""
        import torch
        import torch.distributed as dist
        if not self._distributed() or dist.get_world_size() > 8:
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            nbytes = int(lib.mpu_fusion_mailbox_bytes())
            box = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
            box.zero_()
            hdl = symm.rendezvous(box, group if group is not None else dist.group.WORLD)
            ptrs = [int(p_) for p_ in hdl.buffer_ptrs]
            torch.cuda.synchronize()
            dist.barrier()
        except Exception as e:  # noqa: BLE001 - fall back to NCCL, say why
            print("FusionModel: peer exchange unavailable (%s); using NCCL all-reduce per batch" % (e,))
            return False
        self._peer = dict(box=box, hdl=hdl, ptrs=(ctypes.c_void_p * len(ptrs))(*ptrs), world=dist.get_world_size(),
                          rank=dist.get_rank(), seq=1)
        return True

    def train_on_batch(self, X, y, all_reduce=True, index=None, loss_out=None):
        """One Adam step on a batch of points; X [N,V,C] f32 tensor (device), y [N] uint8 tensor.  With `index`
        (device int64 [n]) the batch is rows index[i] of X / y - a slice of a shuffled epoch, gathered inside the
        kernel.  Single process: ONE launch (mpu_fusion_train_step: gradient sums, the last block applies Adam).
        Several ranks: local sums, SUM all-reduce of the 36 doubles and the point count, identical Adam on every rank.
        Returns the mean dice loss of the GLOBAL batch as a 0-d device tensor (float64)."""
        import torch
        n = int(index.shape[0]) if index is not None else int(X.shape[0])
        self.iterations += 1
        if not (all_reduce and self._distributed()):
            if not hasattr(self, "_counter"):
                self._counter = torch.zeros(int(lib.mpu_fusion_scratch_bytes()), dtype=torch.uint8, device=self.device)
                self._loss1 = torch.zeros(1, dtype=torch.float64, device=self.device)
                self._accum.zero_()
            out = loss_out if loss_out is not None else self._loss1
            check(lib.mpu_fusion_train_step(_C.ptr(X), _C.ptr(y), _C.ptr(index), ctypes.c_longlong(n), self.n_inputs,
                                            self.n_classes, _C.ptr(self.W), _C.ptr(self.b), _C.ptr(self._m),
                                            _C.ptr(self._v), _C.ptr(self._accum), _C.ptr(self._counter), _C.ptr(out),
                                            ctypes.c_float(self.reg), ctypes.c_float(self.lr),
                                            ctypes.c_float(self.beta_1), ctypes.c_float(self.beta_2),
                                            ctypes.c_float(self.epsilon), int(self.iterations), _C.current_stream()),
                  "mpu_fusion_train_step")
            return out[0]
        self._accum.zero_()
        check(lib.mpu_fusion_grad_indexed(_C.ptr(X), _C.ptr(y), _C.ptr(index), ctypes.c_longlong(n), self.n_inputs,
                                          self.n_classes, _C.ptr(self.W), _C.ptr(self.b), _C.ptr(self._accum),
                                          _C.current_stream()), "mpu_fusion_grad_indexed")
        if not hasattr(self, "_red"):
            self._red = torch.zeros(self._accum.numel() + 1, dtype=torch.float64, device=self.device)
        self._red[:-1].copy_(self._accum)
        self._red[-1] = float(n)
        torch.distributed.all_reduce(self._red)          # gradient sums, loss sum and point count in one message
        check(lib.mpu_fusion_adam(_C.ptr(self.W), _C.ptr(self.b), _C.ptr(self._m), _C.ptr(self._v),
                                  _C.ptr(self._red), ctypes.c_double(0.0), self.n_inputs,
                                  self.n_classes, ctypes.c_float(self.reg), ctypes.c_float(self.lr),
                                  ctypes.c_float(self.beta_1), ctypes.c_float(self.beta_2),
                                  ctypes.c_float(self.epsilon), int(self.iterations), _C.current_stream()),
              "mpu_fusion_adam")
        loss = self._red[-2] / self._red[-1]
        if loss_out is not None:
            loss_out.copy_(loss.reshape(1))
        return loss

    def evaluate(self, X, y, batch_size=2 ** 20):
        """Mean dice loss (without regulariser) over a point set, no update (validation_split of the reference's fit)."""
        import torch
        acc = torch.zeros_like(self._accum)
        if int(X.shape[0]) > 0:
            check(lib.mpu_fusion_grad_indexed(_C.ptr(X), _C.ptr(y), _C.ptr(None), ctypes.c_longlong(int(X.shape[0])),
                                              self.n_inputs, self.n_classes, _C.ptr(self.W), _C.ptr(self.b), _C.ptr(acc),
                                              _C.current_stream()), "mpu_fusion_grad_indexed")
        tot = torch.stack([acc[-1], torch.tensor(float(X.shape[0]), dtype=torch.float64, device=self.device)])
        if self._distributed():
            torch.distributed.all_reduce(tot)
        return float((tot[0] / tot[1]).item())

    def fit(self, X, y, batch_size=2 ** 17, epochs=30, verbose=1, shuffle=True, index=None, steps_per_epoch=None, **kw):
        """Epoch loop over device-resident points (bin/train_fusion.py:196-213).  Every epoch draws a permutation and
        walks it in slices: the kernel gathers the rows, X is never copied.  `index` restricts the epoch to a subset of
        the rows (the training part of a validation split); `steps_per_epoch` fixes the number of batches (ranks with
        different point counts must run the same number of collectives: short ranks wrap around)."""
        import torch
        rows = index if index is not None else None
        n = int(rows.shape[0]) if rows is not None else int(X.shape[0])
        nb = steps_per_epoch or (n + batch_size - 1) // batch_size
        hist = []
        for ep in range(epochs):
            perm = torch.randperm(n, device=X.device) if shuffle else torch.arange(n, device=X.device)
            if rows is not None:
                perm = rows[perm]
            if steps_per_epoch and n > 0 and nb * batch_size > n:  # short rank: wrap around to `nb` full batches
                perm = perm.repeat((nb * batch_size + n - 1) // n)[:nb * batch_size]
            losses = torch.zeros(nb, dtype=torch.float64, device=X.device)
            if self._distributed() and getattr(self, "_peer", None) is not None and steps_per_epoch:
                # fused compute + peer-memory exchange: the whole epoch in one C call, one launch per batch
                if not hasattr(self, "_counter"):
                    self._counter = torch.zeros(int(lib.mpu_fusion_scratch_bytes()), dtype=torch.uint8, device=self.device)
                    self._loss1 = torch.zeros(1, dtype=torch.float64, device=self.device)
                    self._accum.zero_()
                pr = self._peer
                perm = perm.contiguous() if n > 0 else torch.zeros(1, dtype=torch.int64, device=X.device)
                check(lib.mpu_fusion_train_epoch_peer(
                    _C.ptr(X), _C.ptr(y), _C.ptr(perm), ctypes.c_longlong(n), ctypes.c_longlong(int(batch_size)),
                    ctypes.c_longlong(int(nb)), self.n_inputs, self.n_classes, _C.ptr(self.W), _C.ptr(self.b),
                    _C.ptr(self._m), _C.ptr(self._v), _C.ptr(self._accum), _C.ptr(self._counter), _C.ptr(losses),
                    ctypes.c_float(self.reg), ctypes.c_float(self.lr), ctypes.c_float(self.beta_1),
                    ctypes.c_float(self.beta_2), ctypes.c_float(self.epsilon), int(self.iterations + 1), pr["ptrs"],
                    pr["world"], pr["rank"], ctypes.c_ulonglong(pr["seq"]), _C.current_stream()),
                    "mpu_fusion_train_epoch_peer")
                pr["seq"] += nb
                self.iterations += nb
                hist.append(float(losses.mean().item()))
                if self.stop_training:
                    break
                continue
            if not self._distributed() and n > 0:
                # the whole epoch in one C call: one fused launch per batch, no host work between them
                if not hasattr(self, "_counter"):
                    self._counter = torch.zeros(int(lib.mpu_fusion_scratch_bytes()), dtype=torch.uint8, device=self.device)
                    self._loss1 = torch.zeros(1, dtype=torch.float64, device=self.device)
                    self._accum.zero_()
                perm = perm[:nb * batch_size].contiguous()
                check(lib.mpu_fusion_train_epoch(_C.ptr(X), _C.ptr(y), _C.ptr(perm), ctypes.c_longlong(int(perm.shape[0])),
                                                 ctypes.c_longlong(int(batch_size)), self.n_inputs, self.n_classes,
                                                 _C.ptr(self.W), _C.ptr(self.b), _C.ptr(self._m), _C.ptr(self._v),
                                                 _C.ptr(self._accum), _C.ptr(self._counter), _C.ptr(losses),
                                                 ctypes.c_float(self.reg), ctypes.c_float(self.lr),
                                                 ctypes.c_float(self.beta_1), ctypes.c_float(self.beta_2),
                                                 ctypes.c_float(self.epsilon), int(self.iterations + 1),
                                                 _C.current_stream()), "mpu_fusion_train_epoch")
                self.iterations += (int(perm.shape[0]) + batch_size - 1) // batch_size
                hist.append(float(losses.mean().item()))
                if self.stop_training:
                    break
                continue
            for k in range(nb):
                idx = perm[k * batch_size:(k + 1) * batch_size]
                if idx.numel() == 0 and not self._distributed():
                    break  # (a rank with no points still joins every all-reduce)
                self.train_on_batch(X, y, index=idx, loss_out=losses[k:k + 1])
            hist.append(float(losses.mean().item()))
            if self.stop_training:
                break
        return hist
