"""ORACLE tooling: a minimal stand-in for the part of `tensorflow.keras` that the reference's model files use, so that
the UNMODIFIED reference graph builders (`/root/reference/mpunet/models/unet.py:114-216`, `fusion_model.py:14-75`) can be
executed in this container, where TensorFlow 2.3.2 cannot be installed.

What this pins and what it does not:
  * PINNED to the reference's own source: the graph - layer order, names, filter counts `int(filters * cf)`, kernel
    sizes, which tensors are concatenated and in which order, cropping, the receptive-field / label-crop attributes,
    `count_params()`.  The reference code builds the graph; this module only records and executes it.
  * NOT pinned (still restated from the Keras 2.x documentation, independently of oracle/unet.py - plain numpy here,
    torch there): the arithmetic INSIDE a layer (Conv2D SAME padding k=2 -> bottom/right, BatchNormalization eps=1e-3
    with moving statistics at inference, MaxPooling2D 2x2 VALID, UpSampling2D nearest, softmax over the last axis).

Functional-API subset: Input, Conv2D, BatchNormalization, MaxPooling2D, UpSampling2D, Cropping2D, Concatenate, Reshape,
Model(inputs, outputs) with .layers (creation order, like Keras), .get_layer, .count_params, .predict, .output;
regularizers.l2; a user-subclassable Layer (build / add_weight / call) and eager numpy versions of the elementary
TensorFlow ops (reduce_sum / mean / max, square, multiply, where, one_hot, softmax, reciprocal ...) that
`fusion_model.py` and `evaluate/loss_functions.py:23-31,207-246` are written in, so that the reference's FusionLayer.call,
its weight regulariser and sparse_generalized_dice_loss run as written.  Only `install()`ed by tests / oracle/make_golden.py; the GPU box has no /root/reference.
"""
import sys
import types

import numpy as np


class _Shape(object):
    def __init__(self, dims):
        self.dims = list(dims)

    def as_list(self):
        return list(self.dims)


class Node(object):
    """Symbolic tensor: shape [None, H, W, C] + the layer call that produces it."""

    def __init__(self, shape, layer=None, inputs=()):
        self.shape = tuple(shape)
        self.layer = layer
        self.inputs = tuple(inputs)

    def get_shape(self):
        return _Shape(self.shape)

    def __repr__(self):
        return "<Node %s from %s>" % (self.shape, getattr(self.layer, "name", None))


_CREATED = []          # layers in creation order since the last Input()
_NAME_COUNTS = {}


def _auto_name(prefix):
    n = _NAME_COUNTS.get(prefix, 0)
    _NAME_COUNTS[prefix] = n + 1
    return prefix if n == 0 else "%s_%d" % (prefix, n)


def _pair(v):
    return (v, v) if isinstance(v, (int, np.integer)) else tuple(int(a) for a in v)


class Layer(object):
    auto_prefix = "layer"

    def __init__(self, name=None, **kw):
        self.name = name or _auto_name(self.auto_prefix)
        self.weights = {}
        self.input = None
        self.output = None
        _CREATED.append(self)

    def __call__(self, x):
        self.input = x
        ins = tuple(x) if isinstance(x, (list, tuple)) else (x,)
        self.build([i.shape for i in ins])
        self.output = Node(self.out_shape([i.shape for i in ins]), self, ins)
        return self.output

    def build(self, in_shapes):
        pass

    def count_params(self):
        """Keras semantics: ALL weights of the layer, trainable or not (BatchNormalization's moving statistics count)."""
        return int(sum(v.size for v in self.weights.values()))

    def trainable_count(self):
        return int(sum(v.size for k, v in self.weights.items() if k not in ("moving_mean", "moving_variance")))


class InputLayer(Layer):
    auto_prefix = "input"

    def run(self, xs):
        return xs[0]


def Input(shape=None, **kw):
    del _CREATED[:]
    _NAME_COUNTS.clear()
    lay = InputLayer(name=kw.get("name"))
    node = Node((None,) + tuple(shape), lay, ())
    lay.input = node          # conv_arithmetics.py:62 reads layers[0].input.get_shape()
    lay.output = node
    return node


def _activation(name):
    if name in (None, "linear"):
        return lambda z: z
    if name == "relu":
        return lambda z: np.maximum(z, 0)
    if name == "softmax":
        def softmax(z):
            e = np.exp(z - z.max(axis=-1, keepdims=True))
            return e / e.sum(axis=-1, keepdims=True)
        return softmax
    if name == "sigmoid":
        return lambda z: 1.0 / (1.0 + np.exp(-z))
    raise NotImplementedError("activation %r" % (name,))


class Conv2D(Layer):
    auto_prefix = "conv2d"

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", activation=None, use_bias=True,
                 dilation_rate=(1, 1), kernel_regularizer=None, name=None, **kw):
        super().__init__(name)
        self.filters = int(filters)
        self.kernel_size = _pair(kernel_size)
        self.strides = _pair(strides)
        self.dilation_rate = _pair(dilation_rate)
        self.padding = padding.lower()
        self.activation_name = activation
        self.activation = _activation(activation)
        self.use_bias = use_bias
        self.kernel_regularizer = kernel_regularizer
        assert self.strides == (1, 1) and self.dilation_rate == (1, 1), "shim: stride / dilation 1 only"

    def build(self, in_shapes):
        cin = in_shapes[0][-1]
        kh, kw = self.kernel_size
        limit = np.sqrt(6.0 / (kh * kw * cin + kh * kw * self.filters))        # glorot_uniform
        self.weights = {"kernel": np.random.uniform(-limit, limit, (kh, kw, cin, self.filters)).astype(np.float32)}
        if self.use_bias:
            self.weights["bias"] = np.zeros(self.filters, np.float32)

    def out_shape(self, s):
        n, h, w, _ = s[0]
        if self.padding == "same":
            return (n, h, w, self.filters)
        return (n, h - self.kernel_size[0] + 1, w - self.kernel_size[1] + 1, self.filters)

    def run(self, xs):
        x = xs[0].astype(np.float64)
        k = self.weights["kernel"].astype(np.float64)
        kh, kw = self.kernel_size
        if self.padding == "same":  # TF: total pad k-1, (k-1)//2 before, the rest after
            pt, pl = (kh - 1) // 2, (kw - 1) // 2
            x = np.pad(x, ((0, 0), (pt, kh - 1 - pt), (pl, kw - 1 - pl), (0, 0)))
        H, W = x.shape[1] - kh + 1, x.shape[2] - kw + 1
        z = np.zeros((x.shape[0], H, W, self.filters))
        for dy in range(kh):        # cross-correlation, HWIO kernel
            for dx in range(kw):
                z += x[:, dy:dy + H, dx:dx + W, :] @ k[dy, dx]
        if self.use_bias:
            z += self.weights["bias"].astype(np.float64)
        return self.activation(z)


class BatchNormalization(Layer):
    auto_prefix = "batch_normalization"

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, name=None, **kw):
        super().__init__(name)
        assert axis in (-1, 3)
        self.momentum, self.epsilon = momentum, epsilon

    def build(self, in_shapes):
        c = in_shapes[0][-1]
        self.weights = {"gamma": np.ones(c, np.float32), "beta": np.zeros(c, np.float32),
                        "moving_mean": np.zeros(c, np.float32), "moving_variance": np.ones(c, np.float32)}

    def out_shape(self, s):
        return s[0]

    def run(self, xs):  # inference: moving statistics
        w = {k: v.astype(np.float64) for k, v in self.weights.items()}
        return (xs[0] - w["moving_mean"]) / np.sqrt(w["moving_variance"] + self.epsilon) * w["gamma"] + w["beta"]


class MaxPooling2D(Layer):
    auto_prefix = "max_pooling2d"

    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", name=None, **kw):
        super().__init__(name)
        self.pool_size = _pair(pool_size)
        self.strides = _pair(strides) if strides is not None else self.pool_size
        assert self.pool_size == (2, 2) and self.strides == (2, 2) and padding == "valid"

    def out_shape(self, s):
        n, h, w, c = s[0]
        return (n, h // 2, w // 2, c)

    def run(self, xs):
        x = xs[0]
        n, h, w, c = x.shape
        x = x[:, :h // 2 * 2, :w // 2 * 2, :].reshape(n, h // 2, 2, w // 2, 2, c)
        return x.max(axis=(2, 4))


class UpSampling2D(Layer):
    auto_prefix = "up_sampling2d"

    def __init__(self, size=(2, 2), interpolation="nearest", name=None, **kw):
        super().__init__(name)
        self.size = _pair(size)
        assert interpolation == "nearest"

    def out_shape(self, s):
        n, h, w, c = s[0]
        return (n, h * self.size[0], w * self.size[1], c)

    def run(self, xs):
        return np.repeat(np.repeat(xs[0], self.size[0], axis=1), self.size[1], axis=2)


class Cropping2D(Layer):
    auto_prefix = "cropping2d"

    def __init__(self, cropping=((0, 0), (0, 0)), name=None, **kw):
        super().__init__(name)
        self.cropping = tuple(tuple(int(a) for a in c) for c in np.asarray(cropping))

    def out_shape(self, s):
        n, h, w, c = s[0]
        (t, b), (l, r) = self.cropping
        return (n, h - t - b, w - l - r, c)

    def run(self, xs):
        (t, b), (l, r) = self.cropping
        x = xs[0]
        return x[:, t:x.shape[1] - b, l:x.shape[2] - r, :]


class Concatenate(Layer):
    auto_prefix = "concatenate"

    def __init__(self, axis=-1, name=None, **kw):
        super().__init__(name)
        assert axis in (-1, 3)

    def out_shape(self, s):
        return s[0][:3] + (sum(a[-1] for a in s),)

    def run(self, xs):
        return np.concatenate(xs, axis=-1)


class Reshape(Layer):
    auto_prefix = "reshape"

    def __init__(self, target_shape, name=None, **kw):
        super().__init__(name)
        self.target_shape = tuple(int(a) for a in target_shape)

    def out_shape(self, s):
        return (s[0][0],) + self.target_shape

    def run(self, xs):
        return xs[0].reshape((xs[0].shape[0],) + self.target_shape)


class Model(object):
    """Model(inputs, outputs): `.layers` lists every layer created since Input() in creation order (what Keras'
    functional Model reports for a graph built top to bottom, which is how the reference builds it)."""

    def __init__(self, inputs=None, outputs=None, **kw):
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        self.layers = list(_CREATED)
        self.output = self.outputs[0]
        self.input = self.inputs[0]

    def get_layer(self, name):
        for lay in self.layers:
            if lay.name == name:
                return lay
        raise ValueError("no layer %r" % name)

    def count_params(self):
        return int(sum(lay.count_params() for lay in self.layers))

    def trainable_count(self):
        return int(sum(lay.trainable_count() for lay in self.layers))

    def predict(self, x, batch_size=None, verbose=0):
        cache = {id(self.inputs[0]): np.asarray(x, dtype=np.float64)}

        def ev(node):
            if id(node) not in cache:
                cache[id(node)] = node.layer.run([ev(i) for i in node.inputs])
            return cache[id(node)]
        return ev(self.outputs[0]).astype(np.float32)

    predict_on_batch = predict


class _L2(object):
    def __init__(self, l2=0.01):
        self.l2 = l2


# ---- user-subclassable Layer + the handful of TensorFlow ops the fusion model and its loss use ------------------------
class T(np.ndarray):
    """ndarray with the two TensorFlow tensor methods the reference calls (get_shape / shape behave like a TensorShape
    for indexing and len)."""

    def get_shape(self):
        return tuple(self.shape)


def _t(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(T)


class KerasLayer(Layer):
    """tensorflow.keras.layers.Layer for subclasses that define build(input_shape) / call(x) (fusion_model.py:14-43)."""
    auto_prefix = "layer"

    def __init__(self, name=None, **kw):
        super().__init__(name)
        self.built = False
        self.regularizers = {}

    def add_weight(self, name, shape, initializer=None, trainable=True, regularizer=None, **kw):
        arr = _t(initializer(shape) if initializer is not None else np.zeros(shape, np.float32))
        self.weights[name] = arr
        if regularizer is not None:
            self.regularizers[name] = regularizer
        return arr

    def build(self, input_shape):
        self.built = True

    def __call__(self, x):
        self.input = x
        if not self.built:
            self.build(tuple(x.shape))
        shp = self.compute_output_shape(tuple(x.shape))
        self.output = Node(tuple(shp), self, (x,))
        return self.output

    def run(self, xs):
        return np.asarray(self.call(_t(xs[0], np.float32)))

    def regularization_losses(self):
        return [float(fn(self.weights[k])) for k, fn in self.regularizers.items()]


def _constant(value):
    return lambda shape: np.full(tuple(int(a) for a in shape), value, dtype=np.float32)


class _Reduction(object):
    NONE, SUM, SUM_OVER_BATCH_SIZE, AUTO = "none", "sum", "sum_over_batch_size", "auto"


class _LossFunctionWrapper(object):
    """tensorflow.python.keras.losses.LossFunctionWrapper: fn(y_true, y_pred, **kwargs) per sample, then the reduction
    (SUM_OVER_BATCH_SIZE = mean over all per-sample values; an optional sample_weight multiplies them first)."""

    def __init__(self, fn, reduction=_Reduction.AUTO, name=None, **kwargs):
        self.fn, self.reduction, self.name, self._fn_kwargs = fn, reduction, name, kwargs

    def __call__(self, y_true, y_pred, sample_weight=None):
        per = np.asarray(self.fn(_t(y_true), _t(y_pred), **self._fn_kwargs))
        if sample_weight is not None:
            per = per * np.asarray(sample_weight).reshape((-1,) + (1,) * (per.ndim - 1))
        if self.reduction == _Reduction.NONE:
            return per
        if self.reduction == _Reduction.SUM:
            return per.sum()
        return per.sum() / per.size


def _tf_ops(tf):
    """Eager numpy implementations of the elementary ops used by mpunet/models/fusion_model.py and
    mpunet/evaluate/loss_functions.py:23-31,207-246 (reductions, elementwise math, one_hot, softmax, where)."""
    def axis_of(axis):
        if axis is None:
            return None
        if isinstance(axis, (int, np.integer)):
            return int(axis)
        return tuple(int(a) for a in axis)      # a range / list; () reduces nothing, as in TensorFlow

    tf.float32, tf.float64, tf.uint8, tf.int32, tf.int64 = np.float32, np.float64, np.uint8, np.int32, np.int64
    tf.convert_to_tensor = lambda x, dtype=None: _t(x, dtype)
    tf.cast = lambda x, dtype: _t(np.asarray(x).astype(dtype))
    tf.size = lambda x: np.asarray(x).size
    tf.shape = lambda x: np.asarray(np.asarray(x).shape)
    tf.reshape = lambda x, shape: _t(np.reshape(np.asarray(x), tuple(int(a) for a in np.asarray(shape))))
    tf.equal = lambda a, b: np.asarray(a) == np.asarray(b)
    tf.cond = lambda pred, true_fn, false_fn: true_fn() if bool(np.all(pred)) else false_fn()
    tf.square = lambda x: _t(np.square(x))
    tf.multiply = lambda a, b: _t(np.multiply(a, b))
    tf.ones_like = lambda x: _t(np.ones_like(x))
    tf.zeros_like = lambda x: _t(np.zeros_like(x))
    tf.where = lambda c, a, b: _t(np.where(c, a, b))
    tf.reduce_sum = lambda x, axis=None, keepdims=False: _t(np.sum(x, axis=axis_of(axis), keepdims=keepdims))
    tf.reduce_mean = lambda x, axis=None, keepdims=False: _t(np.mean(x, axis=axis_of(axis), keepdims=keepdims))
    tf.reduce_max = lambda x, axis=None, keepdims=False: _t(np.max(x, axis=axis_of(axis), keepdims=keepdims))

    def one_hot(indices, depth, dtype=np.float32):
        idx = np.asarray(indices).astype(np.int64)
        out = np.zeros(idx.shape + (int(depth),), dtype=dtype)
        np.put_along_axis(out, idx[..., None], 1, axis=-1)
        return _t(out)
    tf.one_hot = one_hot

    def softmax(z, axis=-1):
        z = np.asarray(z)
        e = np.exp(z - z.max(axis=axis, keepdims=True))
        return _t(e / e.sum(axis=axis, keepdims=True))
    tf.nn = types.ModuleType("tensorflow.nn")
    tf.nn.softmax = softmax
    tf.math = types.ModuleType("tensorflow.math")
    with np.errstate(divide="ignore"):
        pass
    tf.math.reciprocal = lambda x: _t(np.divide(1.0, np.asarray(x), out=np.full(np.shape(x), np.inf, dtype=np.asarray(x).dtype),
                                               where=np.asarray(x) != 0))
    tf.math.square = tf.square
    tf.math.is_inf = lambda x: np.isinf(np.asarray(x))


def install():
    """Registers the stand-in as tensorflow.keras.{models,layers,regularizers} (replacing ref_shim's bare stub, whose
    keras.utils.Sequence is kept) and restores the numpy aliases the reference still uses (np.int: conv_arithmetics.py:23)."""
    from . import ref_shim
    ref_shim.install()
    tf = sys.modules["tensorflow"]
    keras = sys.modules["tensorflow.keras"]
    layers = types.ModuleType("tensorflow.keras.layers")
    for cls in (Conv2D, BatchNormalization, MaxPooling2D, UpSampling2D, Cropping2D, Concatenate, Reshape, Layer):
        setattr(layers, cls.__name__, cls)
    layers.Input = Input
    models = types.ModuleType("tensorflow.keras.models")
    models.Model = Model
    regularizers = types.ModuleType("tensorflow.keras.regularizers")
    regularizers.l2 = _L2
    layers.Layer = KerasLayer
    # any other layer name (Cropping3D, Conv3D ... of the model families outside the hot path) resolves to an inert
    # placeholder class, so that reference modules which merely IMPORT those families can be loaded (PEP 562)
    layers.__getattr__ = lambda name: type(name, (Layer,), {})
    initializers = types.ModuleType("tensorflow.keras.initializers")
    initializers.constant = _constant
    losses = types.ModuleType("tensorflow.keras.losses")
    losses.Reduction = _Reduction
    for name, mod in (("layers", layers), ("models", models), ("regularizers", regularizers),
                      ("initializers", initializers), ("losses", losses)):
        sys.modules["tensorflow.keras." + name] = mod
        setattr(keras, name, mod)
    # inert stand-ins for the keras sub-modules that reference modules outside the arithmetic path import at load time
    # (callbacks, optimizers, metrics, activations, backend ...): any attribute resolves to a placeholder class
    for name in ("callbacks", "optimizers", "metrics", "activations", "backend", "utils"):
        mod = sys.modules.get("tensorflow.keras." + name) or types.ModuleType("tensorflow.keras." + name)
        if "__getattr__" not in mod.__dict__:
            mod.__getattr__ = lambda attr, _m=name: type(attr, (object,), {})
        sys.modules["tensorflow.keras." + name] = mod
        setattr(keras, name, mod)
    if "__getattr__" not in losses.__dict__:
        losses.__getattr__ = lambda attr: type(attr, (object,), {})
    tf.keras = keras
    _tf_ops(tf)
    for name in ("tensorflow.python", "tensorflow.python.keras"):
        sys.modules.setdefault(name, types.ModuleType(name))
    pk_losses = types.ModuleType("tensorflow.python.keras.losses")
    pk_losses.LossFunctionWrapper = _LossFunctionWrapper
    sys.modules["tensorflow.python.keras.losses"] = pk_losses
    if not hasattr(np, "int"):
        np.int = int  # removed in numpy 1.24; the reference pins an older numpy


def _load_reference_file(rel, modname):
    import importlib.util
    import os
    from . import ref_shim
    install()
    spec = importlib.util.spec_from_file_location(modname, os.path.join(ref_shim.REF_ROOT, *rel.split("/")))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_fusion_modules():
    """(loss_functions module, fusion_model module) of the reference, loaded from their own files: the package
    __init__s of mpunet.evaluate / mpunet.models would pull every metric and model family."""
    lf = _load_reference_file("mpunet/evaluate/loss_functions.py", "_ref_loss_functions")
    pkg = types.ModuleType("mpunet.evaluate")
    pkg.__path__ = []
    sys.modules.setdefault("mpunet.evaluate", pkg)
    sys.modules["mpunet.evaluate.loss_functions"] = lf
    fm = _load_reference_file("mpunet/models/fusion_model.py", "_ref_fusion_model")
    return lf, fm


def reference_unet_class():
    """The reference's UNet class, loaded from its own file (mpunet/models/__init__.py would pull every model family)."""
    import importlib.util
    import os
    from . import ref_shim
    install()
    path = os.path.join(ref_shim.REF_ROOT, "mpunet", "models", "unet.py")
    spec = importlib.util.spec_from_file_location("_ref_models_unet", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.UNet
