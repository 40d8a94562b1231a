"""Host-side mirror of mpunet.models.unet.UNet (mpunet/models/unet.py:20-251) on the B200 engine.

Same constructor keywords as the reference class (unknown ones are swallowed by **kwargs exactly like
unet.py:41), same attributes the callers use (`img_shape`, `n_classes`, `label_crop`, `layers`,
`count_params()`), and the Keras methods the reference's Trainer / predict scripts call: `predict`,
`predict_on_batch`, `train_on_batch`, `compile`, `fit`, `load_weights`, `save_weights`.
All arithmetic happens in libmpunet_b200.so (tcgen05 GEMMs + CUDA-core BN/pool/head kernels); torch
tensors are only the device-memory containers.  There is no CPU path.
"""
import ctypes
import math
import os

import numpy as np

from .. import _C
from .._C import lib, check


class MpuUNetConfig(ctypes.Structure):
    _fields_ = [("H", ctypes.c_int), ("W", ctypes.c_int), ("n_channels", ctypes.c_int),
                ("n_classes", ctypes.c_int), ("depth", ctypes.c_int), ("filters", ctypes.c_int * 8),
                ("max_batch", ctypes.c_int), ("training", ctypes.c_int), ("bn_eps", ctypes.c_float),
                ("bn_momentum", ctypes.c_float)]


class MpuLayerInfo(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 64), ("kind", ctypes.c_int), ("ksize", ctypes.c_int),
                ("cin", ctypes.c_int), ("cout", ctypes.c_int), ("k_phys", ctypes.c_int),
                ("co_phys", ctypes.c_int), ("c0_phys", ctypes.c_int), ("off0", ctypes.c_longlong),
                ("off1", ctypes.c_longlong), ("off2", ctypes.c_longlong), ("off3", ctypes.c_longlong)]


class _Layer(object):
    """Minimal stand-in for a Keras layer: name + get_weights/set_weights in Keras layouts
    (what mpunet/utils/utils.py:189-241 `set_bias_weights` touches on the output layer)."""

    def __init__(self, model, info):
        self._m = model
        self.info = info
        self.name = info["name"]
        self.activation = (lambda x: x) if info["kind"] == 0 else None
        if info["kind"] == 0 and info["name"] == "conv2d":
            self.activation.__name__ = model.out_activation

    def get_weights(self):
        return self._m._get_layer_weights(self.info)

    def set_weights(self, ws):
        self._m._set_layer_weights(self.info, ws)


class _Optimizer(object):
    def __init__(self, lr=5e-5, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **kw):
        self.lr = float(kw.get("learning_rate", lr))
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.iterations = 0


def unet_filters(depth, complexity_factor, init_filters=64):
    cf = np.sqrt(complexity_factor)
    return [int(init_filters * 2 ** i * cf) for i in range(depth + 1)]


class UNet(object):
    def __init__(self, n_classes, img_rows=None, img_cols=None, dim=None, n_channels=1, depth=4,
                 out_activation="softmax", activation="relu", kernel_size=3, padding="same",
                 complexity_factor=1, flatten_output=False, l2_reg=None, logger=None,
                 max_batch=32, training=True, seed=None, device=None, **kwargs):
        import torch
        if not ((img_rows and img_cols) or dim):
            raise ValueError("Must specify either img_rows and img_col or dim")
        if dim:
            img_rows, img_cols = dim, dim
        if out_activation != "softmax" or activation != "relu" or kernel_size != 3 or padding != "same":
            raise NotImplementedError("B200 UNet implements the reference defaults only: relu convs, "
                                      "3x3 kernels, 'same' padding, softmax output")
        if l2_reg is not None and not (0.0 <= float(l2_reg) <= 1.0):
            raise ValueError("l2_reg must be a float in [0, 1], got %r" % (l2_reg,))
        if not torch.cuda.is_available():
            raise RuntimeError("multiplanarunet_b200.UNet needs a CUDA device (no CPU fallback)")
        self.logger = logger or print
        self.img_shape = (img_rows, img_cols, n_channels)
        self.n_classes = int(n_classes)
        self.cf = np.sqrt(complexity_factor)
        self.kernel_size, self.activation, self.out_activation = kernel_size, activation, out_activation
        self.l2_reg, self.padding, self.depth = l2_reg, padding, int(depth)
        self.flatten_output = flatten_output
        self.label_crop = np.array([[0, 0], [0, 0]])
        self.stop_training = False
        self.optimizer = _Optimizer()
        self.loss_scale_mode = "sum"
        self.metrics, self.metrics_names = [], ["loss"]
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.max_batch = int(max_batch)
        self.training = bool(training)

        cfg = MpuUNetConfig()
        cfg.H, cfg.W, cfg.n_channels, cfg.n_classes, cfg.depth = img_rows, img_cols, n_channels, n_classes, depth
        for i, f in enumerate(unet_filters(depth, complexity_factor)):
            cfg.filters[i] = f
        cfg.max_batch, cfg.training = self.max_batch, int(self.training)
        cfg.bn_eps, cfg.bn_momentum = 1e-3, 0.99
        self._cfg = cfg
        n_params, n_state, ws = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_longlong()
        check(lib.mpu_unet_sizes(ctypes.byref(cfg), ctypes.byref(n_params), ctypes.byref(n_state),
                                 ctypes.byref(ws)), "mpu_unet_sizes")
        self.n_flat = n_params.value
        with torch.cuda.device(self.device):
            self.params = torch.zeros(self.n_flat, dtype=torch.float32, device=self.device)
            self.bn_state = torch.zeros(n_state.value, dtype=torch.float32, device=self.device)
            if self.training:
                self.grads = torch.zeros_like(self.params)
                self.adam_m = torch.zeros_like(self.params)
                self.adam_v = torch.zeros_like(self.params)
            else:
                self.grads = self.adam_m = self.adam_v = None
            self.workspace = torch.empty(ws.value, dtype=torch.uint8, device=self.device)
            self._loss_dev = torch.zeros(1, dtype=torch.float64, device=self.device)
            self._l2_sumsq = torch.zeros(1, dtype=torch.float64, device=self.device)
            h = ctypes.c_void_p()
            check(lib.mpu_unet_create(ctypes.byref(cfg), _C.ptr(self.params), _C.ptr(self.grads),
                                      _C.ptr(self.adam_m), _C.ptr(self.adam_v), _C.ptr(self.bn_state),
                                      _C.ptr(self.workspace), ctypes.c_longlong(ws.value),
                                      _C.current_stream(), ctypes.byref(h)), "mpu_unet_create")
        self._h = h
        self._infos = []
        for i in range(lib.mpu_unet_num_layers(self._h)):
            li = MpuLayerInfo()
            check(lib.mpu_unet_layer_info(self._h, i, ctypes.byref(li)), "mpu_unet_layer_info")
            self._infos.append({k: (getattr(li, k).decode() if k == "name" else getattr(li, k))
                                for k, _ in MpuLayerInfo._fields_})
        self.layers = [_Layer(self, i) for i in self._ordered_infos()]
        self.init_weights(seed)

    # ------------------------------------------------------------------ layer table / weights
    def _ordered_infos(self):
        """Keras layer order of unet.py (encoder blocks, bottom, up blocks, output conv last)."""
        by = {i["name"]: i for i in self._infos}
        order = []
        for l in range(self.depth):
            order += ["encoder_L%d_conv1" % l, "encoder_L%d_conv2" % l, "encoder_L%d_BN" % l]
        order += ["bottom_conv1", "bottom_conv2", "bottom_BN"]
        for i in range(self.depth):
            order += ["upsample_L%d_conv1" % i, "upsample_L%d_BN1" % i, "upsample_L%d_conv2" % i,
                      "upsample_L%d_conv3" % i, "upsample_L%d_BN2" % i]
        order.append("conv2d")
        return [by[n] for n in order]

    def _kmap(self, info):
        """logical input channel -> physical K index (concat convs keep a gap between the halves)."""
        cin = info["cin"]
        if info["c0_phys"] == info["k_phys"]:
            return np.arange(cin)
        half = cin // 2
        return np.concatenate([np.arange(half), info["c0_phys"] + np.arange(cin - half)])

    def _get_layer_weights(self, info):
        P = self.params
        if info["kind"] == 0:
            k, co, cop, kp = info["ksize"], info["cout"], info["co_phys"], info["k_phys"]
            w = P[info["off0"]:info["off0"] + k * k * cop * kp].view(k * k, cop, kp).cpu().numpy()
            kern = w[:, :co, :][:, :, self._kmap(info)]          # [taps, co, cin]
            kern = kern.transpose(0, 2, 1).reshape(k, k, info["cin"], co).copy()
            bias = P[info["off1"]:info["off1"] + co].cpu().numpy().copy()
            return [kern, bias]
        c = info["cout"]
        S = self.bn_state
        return [P[info["off0"]:info["off0"] + c].cpu().numpy().copy(),
                P[info["off1"]:info["off1"] + c].cpu().numpy().copy(),
                S[info["off2"]:info["off2"] + c].cpu().numpy().copy(),
                S[info["off3"]:info["off3"] + c].cpu().numpy().copy()]

    def _sync(self):
        check(lib.mpu_unet_sync_weights(self._h, _C.current_stream()), "mpu_unet_sync_weights")

    def _set_layer_weights(self, info, ws, sync=True):
        import torch
        P = self.params
        if info["kind"] == 0:
            k, co, cop, kp = info["ksize"], info["cout"], info["co_phys"], info["k_phys"]
            kern = np.asarray(ws[0], dtype=np.float32)
            if kern.shape != (k, k, info["cin"], co):
                raise ValueError("layer %s: kernel shape %s != %s" % (info["name"], kern.shape,
                                                                      (k, k, info["cin"], co)))
            w = np.zeros((k * k, cop, kp), dtype=np.float32)
            tmp = np.zeros((k * k, co, kp), dtype=np.float32)
            tmp[:, :, self._kmap(info)] = kern.reshape(k * k, info["cin"], co).transpose(0, 2, 1)
            w[:, :co, :] = tmp
            P[info["off0"]:info["off0"] + w.size] = torch.from_numpy(w.ravel()).to(self.device)
            b = np.zeros(cop, dtype=np.float32)
            if len(ws) > 1:
                b[:co] = np.asarray(ws[1], dtype=np.float32)
            P[info["off1"]:info["off1"] + cop] = torch.from_numpy(b).to(self.device)
        else:
            c, cp = info["cout"], info["co_phys"]
            S = self.bn_state
            vals = [np.asarray(w, dtype=np.float32) for w in ws]
            pads = [1.0, 0.0, 0.0, 1.0]
            for buf, off, v, pad in ((P, info["off0"], vals[0], pads[0]), (P, info["off1"], vals[1], pads[1]),
                                     (S, info["off2"], vals[2], pads[2]), (S, info["off3"], vals[3], pads[3])):
                full = np.full(cp, pad, dtype=np.float32)
                full[:c] = v
                buf[off:off + cp] = torch.from_numpy(full).to(self.device)
        if sync:
            self._sync()

    def init_weights(self, seed=None):
        """Keras defaults: glorot_uniform kernels, zero biases, BN gamma=1 beta=0 mean=0 var=1."""
        rng = np.random.RandomState(seed)
        for info in self._ordered_infos():
            if info["kind"] == 0:
                k, cin, co = info["ksize"], info["cin"], info["cout"]
                limit = math.sqrt(6.0 / (k * k * cin + k * k * co))
                kern = rng.uniform(-limit, limit, size=(k, k, cin, co)).astype(np.float32)
                self._set_layer_weights(info, [kern, np.zeros(co, np.float32)], sync=False)
            else:
                c = info["cout"]
                self._set_layer_weights(info, [np.ones(c, np.float32), np.zeros(c, np.float32),
                                               np.zeros(c, np.float32), np.ones(c, np.float32)], sync=False)
        self._sync()

    def get_keras_weights(self):
        out = {}
        for info in self._ordered_infos():
            ws = self._get_layer_weights(info)
            if info["kind"] == 0:
                out[info["name"]] = {"kernel": ws[0], "bias": ws[1]}
            else:
                out[info["name"]] = {"gamma": ws[0], "beta": ws[1], "moving_mean": ws[2],
                                     "moving_variance": ws[3]}
        return out

    def set_keras_weights(self, weights):
        for info in self._ordered_infos():
            if info["name"] not in weights:
                continue
            d = weights[info["name"]]
            if info["kind"] == 0:
                self._set_layer_weights(info, [d["kernel"], d["bias"]], sync=False)
            else:
                self._set_layer_weights(info, [d["gamma"], d["beta"], d["moving_mean"], d["moving_variance"]],
                                        sync=False)
        self._sync()

    def get_flat_grads_as_keras(self):
        """Gradients of the last train step, in Keras layouts (tests / debugging)."""
        import torch
        saved_p, saved_s = self.params, self.bn_state
        try:
            self.params = self.grads
            self.bn_state = torch.zeros_like(saved_s)
            out = {}
            for info in self._ordered_infos():
                ws = self._get_layer_weights(info)
                if info["kind"] == 0:
                    out[(info["name"], "kernel")], out[(info["name"], "bias")] = ws[0], ws[1]
                else:
                    out[(info["name"], "gamma")], out[(info["name"], "beta")] = ws[0], ws[1]
            return out
        finally:
            self.params, self.bn_state = saved_p, saved_s

    def count_params(self):
        n = 0
        for info in self._infos:
            if info["kind"] == 0:
                n += info["ksize"] ** 2 * info["cin"] * info["cout"] + info["cout"]
            else:
                n += 4 * info["cout"]
        return n

    def save_weights(self, path, overwrite=True):
        """Weights only, by Keras layer name (as model.save_weights does, bin/train.py:303-317).  `*.h5` / `*.hdf5`
        paths are written in Keras' HDF5 layout by utils/keras_h5.py (no h5py needed), anything else as `.npz`."""
        if not overwrite and os.path.exists(path):
            raise OSError("%s exists" % path)
        weights = self.get_keras_weights()
        if path.endswith((".h5", ".hdf5")):
            from ..utils.keras_h5 import save_keras_weights
            save_keras_weights(path, weights, layer_order=[i["name"] for i in self._ordered_infos()])
            return
        flat = {}
        for name, d in weights.items():
            for k, v in d.items():
                flat["%s/%s" % (name, k)] = v
        with open(path, "wb") as f:
            np.savez(f, **flat)

    def load_weights(self, path, by_name=True):
        """`.npz` archives of this package or Keras HDF5 weight files (load_weights(path, by_name=True),
        models/model_init.py:31,56): layers are matched by name, shapes must agree."""
        if path.endswith((".h5", ".hdf5")):
            from ..utils.keras_h5 import load_keras_weights
            weights = load_keras_weights(path)
        else:
            with np.load(path) as z:
                weights = {}
                for key in z.files:
                    name, k = key.rsplit("/", 1)
                    weights.setdefault(name, {})[k] = z[key]
        known = set(i["name"] for i in self._infos)
        if not known & set(weights):
            raise ValueError("%s holds no layer of this model (found %s)" % (path, sorted(weights)[:5]))
        self.set_keras_weights(weights)

    # ------------------------------------------------------------------ inference
    def _pack(self, xb):
        import torch
        x = torch.as_tensor(np.ascontiguousarray(xb, dtype=np.float32)) if not torch.is_tensor(xb) else xb
        x = x.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        B = x.shape[0]
        if tuple(x.shape[1:]) != tuple(self.img_shape):
            raise ValueError("expected input [B,%d,%d,%d], got %s" % (self.img_shape + (tuple(x.shape),)))
        check(lib.mpu_unet_pack_input(self._h, _C.ptr(x), B, _C.current_stream()), "mpu_unet_pack_input")
        return B

    def predict_on_batch(self, x, bn_training=False, as_numpy=True):
        import torch
        B = self._pack(x)
        H, W, _ = self.img_shape
        probs = torch.empty(B, H, W, self.n_classes, dtype=torch.float32, device=self.device)
        check(lib.mpu_unet_forward(self._h, B, int(bn_training), _C.ptr(probs), _C.current_stream()),
              "mpu_unet_forward")
        if self.flatten_output:
            probs = probs.view(B, H * W, self.n_classes)
        return probs.cpu().numpy() if as_numpy else probs

    def predict(self, X, batch_size=8, verbose=0):
        X = np.asarray(X)
        bs = min(int(batch_size), self.max_batch)
        outs = [self.predict_on_batch(X[i:i + bs]) for i in range(0, X.shape[0], bs)]
        return np.concatenate(outs, axis=0)

    # ------------------------------------------------------------------ training
    def compile(self, optimizer=None, loss=None, metrics=None, **kwargs):
        if optimizer is not None and not isinstance(optimizer, str):
            self.optimizer = optimizer
        return self

    def forward_backward(self, x, y, sample_weight=None, input_packed=False, batch=None):
        """Forward + loss + backward of one batch; gradients stay in self.grads (for all-reduce)."""
        import torch
        B = batch if input_packed else self._pack(x)
        H, W, _ = self.img_shape
        yy = y if torch.is_tensor(y) else torch.as_tensor(np.ascontiguousarray(y).reshape(B, H, W).astype(np.uint8))
        yy = yy.to(self.device, dtype=torch.uint8, non_blocking=True).contiguous()
        sw = None
        if sample_weight is not None:
            sw = torch.as_tensor(np.asarray(sample_weight, dtype=np.float32)) if not torch.is_tensor(
                sample_weight) else sample_weight
            sw = sw.to(self.device, dtype=torch.float32).contiguous()
        gscale = 1.0 if self.loss_scale_mode == "sum" else 1.0 / (B * H * W)
        check(lib.mpu_unet_train_step(self._h, B, _C.ptr(yy), _C.ptr(sw), ctypes.c_float(gscale),
                                      _C.ptr(self._loss_dev), _C.ptr(None), _C.current_stream()),
              "mpu_unet_train_step")
        self._last_B = B
        return self._loss_dev

    def _train_step_adam(self, x, y, sample_weight=None, input_packed=False, batch=None):
        """Single-process step in one C call (mpu_unet_train_step_adam): forward + loss + backward, with the l2 penalty
        and Adam update of every parameter range issued as soon as its gradients are final, overlapped with the rest
        of backward.  Same results as forward_backward() + apply_gradients()."""
        import torch
        B = batch if input_packed else self._pack(x)
        H, W, _ = self.img_shape
        yy = y if torch.is_tensor(y) else torch.as_tensor(np.ascontiguousarray(y).reshape(B, H, W).astype(np.uint8))
        yy = yy.to(self.device, dtype=torch.uint8, non_blocking=True).contiguous()
        sw = None
        if sample_weight is not None:
            sw = sample_weight if torch.is_tensor(sample_weight) else torch.as_tensor(
                np.asarray(sample_weight, dtype=np.float32))
            sw = sw.to(self.device, dtype=torch.float32).contiguous()
        gscale = 1.0 if self.loss_scale_mode == "sum" else 1.0 / (B * H * W)
        o = self.optimizer
        o.iterations += 1
        l2c = 2.0 * float(self.l2_reg) * gscale * B * H * W if self.l2_reg else 0.0
        check(lib.mpu_unet_train_step_adam(self._h, B, _C.ptr(yy), _C.ptr(sw), ctypes.c_float(gscale),
                                           _C.ptr(self._loss_dev), _C.ptr(None), ctypes.c_float(o.lr),
                                           ctypes.c_float(o.beta_1), ctypes.c_float(o.beta_2),
                                           ctypes.c_float(o.epsilon), int(o.iterations), ctypes.c_float(l2c),
                                           _C.ptr(self._l2_sumsq), _C.current_stream()), "mpu_unet_train_step_adam")
        self._last_B = B
        return self._loss_dev

    def forward_backward_overlapped(self, x, y, sample_weight=None, input_packed=False, batch=None, fuse_adam=False):
        """forward_backward with the gradient all-reduce overlapped with backward: parameter ranges whose
        gradients are final after each backward stage are all-reduced asynchronously (NCCL stream) while
        the next stage computes.  Falls back to forward_backward when no process group is initialised."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            if fuse_adam:
                return self._train_step_adam(x, y, sample_weight, input_packed, batch)
            return self.forward_backward(x, y, sample_weight, input_packed, batch)
        B = batch if input_packed else self._pack(x)
        H, W, _ = self.img_shape
        yy = y if torch.is_tensor(y) else torch.as_tensor(np.ascontiguousarray(y).reshape(B, H, W).astype(np.uint8))
        yy = yy.to(self.device, dtype=torch.uint8, non_blocking=True).contiguous()
        sw = None
        if sample_weight is not None:
            sw = sample_weight if torch.is_tensor(sample_weight) else torch.as_tensor(
                np.asarray(sample_weight, dtype=np.float32))
            sw = sw.to(self.device, dtype=torch.float32).contiguous()
        gscale = 1.0 if self.loss_scale_mode == "sum" else 1.0 / (B * H * W)
        st = _C.current_stream()
        check(lib.mpu_unet_train_forward(self._h, B, _C.ptr(yy), _C.ptr(sw), ctypes.c_float(gscale),
                                         _C.ptr(self._loss_dev), _C.ptr(None), st), "mpu_unet_train_forward")
        if not hasattr(self, "_ranges"):
            r = (ctypes.c_longlong * 8)()
            check(lib.mpu_unet_grad_ranges(self._h, r), "mpu_unet_grad_ranges")
            self._ranges = [(r[2 * i], r[2 * i + 1]) for i in range(4)]
        works = []
        for stage in range(3):
            check(lib.mpu_unet_backward_stage(self._h, B, stage, st), "mpu_unet_backward_stage")
            for (a, b) in ([self._ranges[stage]] if stage < 2 else self._ranges[2:]):
                if b > a:
                    works.append(((a, b), dist.all_reduce(self.grads[a:b], async_op=True)))
        self._last_B = B
        if fuse_adam:
            # Adam range by range, each as soon as ITS all-reduce is done: the update of the early (large) ranges
            # overlaps the reduction of the last one
            o = self.optimizer
            o.iterations += 1
            self._l2_begin()
            for k, ((a, b), w) in enumerate(works):
                w.wait()
                self._l2_penalty(a, b)
                check(lib.mpu_unet_adam_range(self._h, ctypes.c_longlong(a), ctypes.c_longlong(b), ctypes.c_float(o.lr),
                                              ctypes.c_float(o.beta_1), ctypes.c_float(o.beta_2),
                                              ctypes.c_float(o.epsilon), int(o.iterations), ctypes.c_float(1.0),
                                              int(k == len(works) - 1), st), "mpu_unet_adam_range")
            return self._loss_dev
        for _, w in works:
            w.wait()
        return self._loss_dev

    # kernel_regularizer=l2(l2_reg) (mpunet/models/unet.py:122-189).  Keras adds the scalar penalty to every element
    # of the unreduced [B, H*W] loss tensor it differentiates (Reduction.NONE), so next to a data gradient
    # scaled by g the penalty's gradient is g * (B*H*W) * 2 * l2_reg * w; the reported (mean) loss gains l2_reg * sum(w^2).
    def _l2_begin(self):
        if self.l2_reg:
            self._l2_sumsq.zero_()

    def _l2_penalty(self, begin, end):
        if not self.l2_reg:
            return
        H, W, _ = self.img_shape
        g = (1.0 if self.loss_scale_mode == "sum" else 1.0 / (self._last_B * H * W)) * self._last_B * H * W
        check(lib.mpu_unet_l2_penalty(self._h, ctypes.c_longlong(begin), ctypes.c_longlong(end),
                                      ctypes.c_float(2.0 * float(self.l2_reg) * g),
                                      _C.ptr(self._l2_sumsq), _C.current_stream()), "mpu_unet_l2_penalty")

    def apply_gradients(self, grad_scale=1.0):
        o = self.optimizer
        o.iterations += 1
        self._l2_begin()
        self._l2_penalty(0, self.grads.numel())  # Adam's grad_scale then applies to data and penalty alike
        check(lib.mpu_unet_adam(self._h, ctypes.c_float(o.lr), ctypes.c_float(o.beta_1),
                                ctypes.c_float(o.beta_2), ctypes.c_float(o.epsilon), int(o.iterations),
                                ctypes.c_float(grad_scale), _C.current_stream()), "mpu_unet_adam")

    def train_on_batch_async(self, x, y, sample_weight=None):
        """One train step (forward + loss + backward, gradient all-reduce overlapped with backward when a process
        group is up, Adam) without a host synchronisation.  x / y / sample_weight may be numpy arrays, pinned host
        tensors or device tensors.  Returns the mean loss as a 0-d device tensor (float64)."""
        loss = self.forward_backward_overlapped(x, y, sample_weight, fuse_adam=True)
        H, W, _ = self.img_shape
        mean = loss[0] / float(self._last_B * H * W)
        return mean + float(self.l2_reg) * self._l2_sumsq[0] if self.l2_reg else mean

    def train_on_batch(self, x, y, sample_weight=None):
        """Keras' model.train_on_batch: the step above, returning the mean loss as a Python float (one 8-byte
        device->host read)."""
        return float(self.train_on_batch_async(x, y, sample_weight).item())

    def fit(self, x, steps_per_epoch=None, epochs=1, callbacks=None, initial_epoch=0, verbose=1,
            train_on_batch=None, sync_stop=None, **kw):
        """Keras-style fit over a generator of (x, y, w) batches with epoch callbacks
        (train/trainer.py:246-257); returns History.history-like {"loss": [...], ...}."""
        from ..train import fit_loop
        return fit_loop(self, x, steps_per_epoch, epochs, callbacks=callbacks, initial_epoch=initial_epoch,
                        train_on_batch=train_on_batch, verbose=verbose, logger=self.logger, sync_stop=sync_stop)

    def reset_metrics(self):
        pass

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib.mpu_unet_destroy(self._h)
                self._h = None
        except Exception:
            pass
