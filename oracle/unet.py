"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement (torch-CPU, fp32) of the reference's 2D U-Net graph and train step:
  * graph ......................... mpunet/models/unet.py:114-216 (encoder :114-134, bottom :136-146,
                                    up path :148-180, 1x1 softmax head :211-214)
  * filter counts ................. int(64 * 2**i * sqrt(complexity_factor)), unet.py:91,120,195
  * loss / optimizer .............. SparseCategoricalCrossentropy(reduction=NONE) x sample_weight,
                                    Adam(lr 5e-5, eps 1e-8): bin/defaults/MultiPlanar/train_hparams.yaml:108,125-126,
                                    train/trainer.py:51-101,246-257

GRAPH PINNED: layer order / names / filter counts / kernel sizes / concatenation order / parameter counts equal the
graph that the reference's own UNet.init_model builds when executed unmodified under oracle/keras_shim.py
(tests/golden/unet_graph_*.npz, tests/test_oracle_golden.py, tests/test_oracle_vs_reference.py).
LAYER ARITHMETIC - PARITY UNPINNED: the arithmetic of these layers lives in TensorFlow 2.3.2 (requirements.txt:10), which
cannot be installed here (no wheel for Python 3.12, no network) and the reference ships no golden
vectors for the network.  Layer semantics restated from the TF/Keras 2.3
documentation: Conv2D NHWC / HWIO cross-correlation, SAME padding = (k-1)//2 before, rest after
(k=2: bottom/right only); ReLU inside the conv; BatchNormalization(axis=-1, momentum=.99, eps=1e-3),
training = batch mean / biased variance (moving variance updated with the unbiased estimate, as the
fused kernel does); MaxPool 2x2/2; UpSampling2D nearest; softmax over channels; sparse CE computed from
the logits (Keras uses the softmax op's input when the prediction comes straight from a softmax);
per-pixel losses are left unreduced and Keras differentiates their SUM (loss_scale="sum").

`emulate_bf16=True` rounds at exactly the points where the CUDA path stores bf16 (conv outputs, BN
outputs, GEMM operand weights, all back-propagated activation gradients), so the two differ only by
fp32 accumulation order.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def filters_for(depth=4, complexity_factor=1.0, init_filters=64):
    cf = np.sqrt(complexity_factor)
    enc = [int(init_filters * 2 ** i * cf) for i in range(depth)]
    bottom = int(init_filters * 2 ** depth * cf)
    up = [int(init_filters * 2 ** (depth - 1 - i) * cf) for i in range(depth)]
    return enc, bottom, up


def layer_specs(n_classes, n_channels=1, depth=4, complexity_factor=1.0):
    """Ordered list of (name, kind, shape...) following the Keras layer names of unet.py:119-179."""
    enc, bottom, up = filters_for(depth, complexity_factor)
    specs = []
    cin = n_channels
    for i, f in enumerate(enc):
        specs.append(("encoder_L%d_conv1" % i, "conv", 3, cin, f))
        specs.append(("encoder_L%d_conv2" % i, "conv", 3, f, f))
        specs.append(("encoder_L%d_BN" % i, "bn", f))
        cin = f
    specs.append(("bottom_conv1", "conv", 3, cin, bottom))
    specs.append(("bottom_conv2", "conv", 3, bottom, bottom))
    specs.append(("bottom_BN", "bn", bottom))
    cin = bottom
    for i, f in enumerate(up):
        specs.append(("upsample_L%d_conv1" % i, "conv", 2, cin, f))
        specs.append(("upsample_L%d_BN1" % i, "bn", f))
        specs.append(("upsample_L%d_conv2" % i, "conv", 3, 2 * f, f))
        specs.append(("upsample_L%d_conv3" % i, "conv", 3, f, f))
        specs.append(("upsample_L%d_BN2" % i, "bn", f))
        cin = f
    specs.append(("conv2d", "conv", 1, cin, n_classes))  # unnamed Keras layer of unet.py:211
    return specs


def init_params(n_classes, n_channels=1, depth=4, complexity_factor=1.0, seed=0, randomize_bn=False):
    """glorot_uniform kernels / zero biases / BN gamma=1 beta=0 mean=0 var=1 (Keras defaults).
    Returns {name: {"kernel": HWIO, "bias": [O]} | {"gamma","beta","moving_mean","moving_variance"}}."""
    rng = np.random.RandomState(seed)
    P = {}
    for spec in layer_specs(n_classes, n_channels, depth, complexity_factor):
        name, kind = spec[0], spec[1]
        if kind == "conv":
            k, cin, cout = spec[2:]
            limit = math.sqrt(6.0 / (k * k * cin + k * k * cout))
            P[name] = {"kernel": rng.uniform(-limit, limit, size=(k, k, cin, cout)).astype(np.float32),
                       "bias": np.zeros(cout, dtype=np.float32)}
        else:
            c = spec[2]
            P[name] = {"gamma": np.ones(c, np.float32), "beta": np.zeros(c, np.float32),
                       "moving_mean": np.zeros(c, np.float32), "moving_variance": np.ones(c, np.float32)}
            if randomize_bn:
                P[name]["gamma"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
                P[name]["beta"] = (0.1 * rng.randn(c)).astype(np.float32)
                P[name]["moving_mean"] = (0.1 * rng.randn(c)).astype(np.float32)
                P[name]["moving_variance"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    return P


def count_params(P):
    return int(sum(v.size for d in P.values() for k, v in d.items()
                   if k in ("kernel", "bias", "gamma", "beta")))


# ---- bf16 emulation helpers ---------------------------------------------------------------------
def _bf16(t):
    return t.to(torch.bfloat16).to(t.dtype)


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _ReluMask(torch.autograd.Function):
    """relu whose BACKWARD mask is given (teacher forcing: the mask of the forced activation, so that an
    activation that is +tiny on the device and -tiny here - or the reverse - does not switch a gradient path)."""

    @staticmethod
    def forward(ctx, pre, mask):
        ctx.save_for_backward(mask)
        return pre.clamp_min(0)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask, None


class _GradTap(torch.autograd.Function):
    """Identity in forward.  In backward it records the gradient the oracle computed at this point
    (`seen[name]`, NHWC) and - teacher forcing of the BACKWARD pass - replaces it by `forced[name]` when
    given, so that every layer's backward arithmetic is checked locally: oracle(layer backward of the
    device's own upstream gradient) against the device's downstream gradient."""

    @staticmethod
    def forward(ctx, x, name, seen, forced):
        ctx.name, ctx.seen, ctx.forced = name, seen, forced
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if ctx.seen is not None:
            ctx.seen[ctx.name] = g.detach().permute(0, 2, 3, 1).numpy().copy()
        if ctx.forced is not None and ctx.name in ctx.forced:
            g = torch.as_tensor(ctx.forced[ctx.name], dtype=g.dtype).permute(0, 3, 1, 2).contiguous()
        return g, None, None, None


def collapse_upconv_weights(w):
    """w [2,2,Cin,Cout] -> dict {(a,b): [((di,dj), W[Cin,Cout]) ...]}: the four sub-pixel phases of
    nearest-2x-upsample followed by the 2x2 SAME conv (unet.py:159-163) as 1/2/2/4-tap convs on the
    low-res grid: out[2i+a, 2j+b] = sum_taps Wc . in[i+di, j+dj]."""
    out = {}
    for a in range(2):
        for b in range(2):
            acc = {}
            for dy in range(2):
                for dx in range(2):
                    key = ((a + dy) >> 1, (b + dx) >> 1)
                    acc[key] = acc[key] + w[dy, dx] if key in acc else w[dy, dx]
            out[(a, b)] = sorted(acc.items())
    return out


class UNetOracle:
    def __init__(self, n_classes, n_channels=1, depth=4, complexity_factor=1.0, params=None, seed=0,
                 dtype=torch.float32):
        self.n_classes, self.n_channels, self.depth = n_classes, n_channels, depth
        self.cf = complexity_factor
        self.dtype = dtype
        P = params if params is not None else init_params(n_classes, n_channels, depth,
                                                          complexity_factor, seed)
        self.P = {n: {k: torch.tensor(v, dtype=dtype) for k, v in d.items()} for n, d in P.items()}
        self.bn_eps = 1e-3
        self.bn_momentum = 0.99

    def trainable(self):
        out = []
        for n, d in self.P.items():
            for k in ("kernel", "bias", "gamma", "beta"):
                if k in d:
                    out.append((n, k, d[k]))
        return out

    # -- layers (NCHW internally) -----------------------------------------------------------------
    def _conv(self, x, name, emu, relu=True, mask=None):
        w = self.P[name]["kernel"]  # HWIO
        b = self.P[name]["bias"]
        k = w.shape[0]
        if emu:
            x = _RoundBwd.apply(x)
            wq = _RoundFwd.apply(w)
        else:
            wq = w
        wt = wq.permute(3, 2, 0, 1)
        if k == 3:
            y = F.conv2d(x, wt, padding=1)
        elif k == 1:
            y = F.conv2d(x, wt)
        else:
            raise ValueError(k)
        y = y + b.view(1, -1, 1, 1)
        if relu:
            y = F.relu(y) if mask is None else _ReluMask.apply(y, mask)
        return _RoundFwd.apply(y) if emu else y

    def _upconv(self, x, name, emu, mask=None):
        """UpSampling2D(2) + Conv2D(k=2, SAME, relu) (unet.py:159-163)."""
        w = self.P[name]["kernel"]
        b = self.P[name]["bias"]
        if not emu:
            up = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
            up = F.pad(up, (0, 1, 0, 1))
            y = F.conv2d(up, w.permute(3, 2, 0, 1)) + b.view(1, -1, 1, 1)
            return F.relu(y) if mask is None else _ReluMask.apply(y, mask)
        x = _RoundBwd.apply(x)
        B, C, h, ww = x.shape
        xp = F.pad(x, (0, 1, 0, 1))
        y = torch.zeros(B, w.shape[3], 2 * h, 2 * ww, dtype=x.dtype)
        phases = collapse_upconv_weights(w)
        for (a, bb), taps in phases.items():
            acc = 0
            for (di, dj), wc in taps:
                wq = _RoundFwd.apply(wc)  # [Cin, Cout]
                acc = acc + torch.einsum("bchw,co->bohw", xp[:, :, di:di + h, dj:dj + ww], wq)
            y[:, :, a::2, bb::2] = acc
        y = y + b.view(1, -1, 1, 1)
        y = F.relu(y) if mask is None else _ReluMask.apply(y, mask)
        return _RoundFwd.apply(y)

    def _bn(self, x, name, training, emu, stats_out=None):
        d = self.P[name]
        if emu:
            x = _RoundBwd.apply(x)
        if training:
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            if stats_out is not None:
                n = x.numel() / x.shape[1]
                stats_out[name] = (mean.detach().clone(), (var * n / max(n - 1, 1)).detach().clone())
        else:
            mean, var = d["moving_mean"], d["moving_variance"]
        scale = d["gamma"] / torch.sqrt(var + self.bn_eps)
        shift = d["beta"] - mean * scale
        y = x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        return _RoundFwd.apply(y) if emu else y

    def logits(self, x_nhwc, training=False, emulate_bf16=False, stats_out=None, capture=None,
               force=None, computed=None, grad_seen=None, force_grad=None):
        """`force` {name: NHWC array}: teacher forcing - the forward VALUE of the named activation is
        replaced by the given one (gradients still flow through the computed expression), so a backward
        comparison is not polluted by forward rounding chaos.
        `computed` (out): the value the oracle computed for each activation BEFORE forcing, i.e. this
        layer's output from the forced inputs - the per-layer forward residual is computed[name] vs force[name].
        `grad_seen` (out) / `force_grad` (in): the same for the backward pass (see _GradTap); extra tap
        "skip_<l>" sits on the skip branch of the concat."""
        emu = emulate_bf16
        taps = grad_seen is not None or force_grad is not None

        def cap(name, t):
            if computed is not None:
                computed[name] = t.detach().permute(0, 2, 3, 1).numpy().copy()
            if force is not None and name in force:
                f = torch.as_tensor(force[name], dtype=t.dtype).permute(0, 3, 1, 2).contiguous()
                t = f + (t - t.detach())  # value exactly f, gradient flows to the computed expression
            if capture is not None:
                capture[name] = t.detach().permute(0, 2, 3, 1).numpy().copy()
            if taps and t.requires_grad:
                t = _GradTap.apply(t, name, grad_seen, force_grad)
            return t

        def fm(name):
            """ReLU backward mask of a forced post-ReLU activation (None when not forced)."""
            if force is None or name not in force:
                return None
            return (torch.as_tensor(force[name]).permute(0, 3, 1, 2) > 0).to(self.dtype).contiguous()

        x = torch.as_tensor(x_nhwc, dtype=self.dtype).permute(0, 3, 1, 2)
        if emu:
            x = _bf16(x)
        skips = []
        for i in range(self.depth):
            x = self._conv(x, "encoder_L%d_conv1" % i, emu, mask=fm("a1_%d" % i))
            x = cap("a1_%d" % i, x)
            x = self._conv(x, "encoder_L%d_conv2" % i, emu, mask=fm("a2_%d" % i))
            x = cap("a2_%d" % i, x)
            x = self._bn(x, "encoder_L%d_BN" % i, training, emu, stats_out)
            x = cap("b_%d" % i, x)
            skips.append(x)
            x = F.max_pool2d(x, 2)
            x = cap("pooled_%d" % i, x)
        x = self._conv(x, "bottom_conv1", emu, mask=fm("a1_%d" % self.depth))
        x = cap("a1_%d" % self.depth, x)
        x = self._conv(x, "bottom_conv2", emu, mask=fm("a2_%d" % self.depth))
        x = cap("a2_%d" % self.depth, x)
        x = self._bn(x, "bottom_BN", training, emu, stats_out)
        x = cap("b_%d" % self.depth, x)
        for i in range(self.depth):
            l = self.depth - 1 - i
            x = self._upconv(x, "upsample_L%d_conv1" % i, emu, mask=fm("u_%d" % l))
            x = cap("u_%d" % l, x)
            x = self._bn(x, "upsample_L%d_BN1" % i, training, emu, stats_out)
            x = cap("bn1_%d" % l, x)
            sk = skips[self.depth - 1 - i]
            if taps and sk.requires_grad:
                sk = _GradTap.apply(sk, "skip_%d" % l, grad_seen, force_grad)
            x = torch.cat([sk, x], dim=1)
            x = self._conv(x, "upsample_L%d_conv2" % i, emu, mask=fm("c2_%d" % l))
            x = cap("c2_%d" % l, x)
            x = self._conv(x, "upsample_L%d_conv3" % i, emu, mask=fm("c3_%d" % l))
            x = cap("c3_%d" % l, x)
            x = self._bn(x, "upsample_L%d_BN2" % i, training, emu, stats_out)
            x = cap("bn2_%d" % l, x)
        if emu:
            x = _RoundBwd.apply(x)
        w = self.P["conv2d"]["kernel"]
        z = F.conv2d(x, w.permute(3, 2, 0, 1)) + self.P["conv2d"]["bias"].view(1, -1, 1, 1)
        return z.permute(0, 2, 3, 1)  # NHWC

    def predict(self, x_nhwc, emulate_bf16=False, training=False, capture=None):
        with torch.no_grad():
            return torch.softmax(self.logits(x_nhwc, training, emulate_bf16, capture=capture), dim=-1).numpy()

    def loss_and_grads(self, x_nhwc, y, sample_weight=None, emulate_bf16=False, loss_scale="sum",
                       force=None, computed=None, grad_seen=None, force_grad=None, capture=None):
        """One training forward/backward.  y [B,H,W] (or [B,HW,1]) integer labels.
        Returns (mean loss, {(layer, param): grad ndarray}, batch BN stats {name: (mean, unbiased var)})."""
        for _, _, t in self.trainable():
            t.requires_grad_(True)
            t.grad = None
        stats = {}
        z = self.logits(x_nhwc, True, emulate_bf16, stats, capture=capture, force=force, computed=computed,
                        grad_seen=grad_seen, force_grad=force_grad)
        B, H, W, C = z.shape
        yy = torch.as_tensor(np.asarray(y).reshape(B, H, W).astype(np.int64))
        logp = torch.log_softmax(z, dim=-1)
        ce = -logp.gather(-1, yy.unsqueeze(-1)).squeeze(-1)
        sw = torch.ones(B, dtype=self.dtype) if sample_weight is None else torch.as_tensor(
            np.asarray(sample_weight), dtype=self.dtype)
        per = ce * sw.view(B, 1, 1)
        total = per.sum() if loss_scale == "sum" else per.mean()
        total.backward()
        grads = {(n, k): t.grad.detach().numpy().copy() for n, k, t in self.trainable()}
        for _, _, t in self.trainable():
            t.requires_grad_(False)
        return float(per.mean().item()), grads, {k: (m.numpy(), v.numpy()) for k, (m, v) in stats.items()}
