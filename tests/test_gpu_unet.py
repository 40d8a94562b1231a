"""GPU parity tests of the U-Net engine against the CPU oracle (oracle/unet.py).

Tolerances (floating point path, bf16 storage / fp32 accumulate):
  * inference: softmax probabilities within 1e-3 max-abs of the oracle that rounds to bf16 at the same
    points, arg-max label maps bit-exact (north_star bar);
  * train step: loss equal to 1e-5 relative; gradients against the teacher-forced oracle (forward values
    pinned to the GPU's own activations so rounding chaos of the random net does not pollute the
    comparison): cosine > 0.999 and max error < 6% of the gradient's max per tensor;
  * BatchNorm moving statistics after one step within 1e-5.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = dict(dim=64, batch=4, cf=0.125, classes=3, channels=1)


@pytest.fixture(scope="module")
def setup():
    import torch
    from multiplanarunet_b200.models import UNet
    from oracle.unet import UNetOracle, init_params
    rng = np.random.RandomState(0)
    P = init_params(CFG["classes"], CFG["channels"], 4, CFG["cf"], seed=1, randomize_bn=True)
    for n, d in P.items():
        if "bias" in d:
            d["bias"] = (0.05 * rng.randn(*d["bias"].shape)).astype(np.float32)
    x = rng.randn(CFG["batch"], CFG["dim"], CFG["dim"], CFG["channels"]).astype(np.float32)
    y = rng.randint(0, CFG["classes"], size=(CFG["batch"], CFG["dim"], CFG["dim"])).astype(np.uint8)
    sw = rng.uniform(0.5, 1.5, size=CFG["batch"]).astype(np.float32)
    model = UNet(n_classes=CFG["classes"], dim=CFG["dim"], n_channels=CFG["channels"],
                 complexity_factor=CFG["cf"], max_batch=CFG["batch"], training=True)
    model.set_keras_weights(P)
    return dict(P=P, x=x, y=y, sw=sw, model=model, oracle=UNetOracle(CFG["classes"], CFG["channels"], 4, CFG["cf"], params=P))


def test_weight_layout_round_trip(setup):
    P2 = setup["model"].get_keras_weights()
    for n in setup["P"]:
        for k in setup["P"][n]:
            assert np.array_equal(setup["P"][n][k], P2[n][k]), (n, k)
    from oracle.unet import count_params
    assert setup["model"].count_params() == count_params(setup["P"]) + sum(
        2 * d["moving_mean"].size for d in setup["P"].values() if "moving_mean" in d)


def test_inference_probabilities_and_labels(setup):
    got = setup["model"].predict_on_batch(setup["x"])
    ref = setup["oracle"].predict(setup["x"], emulate_bf16=True)
    ref32 = setup["oracle"].predict(setup["x"], emulate_bf16=False)
    assert np.abs(got - ref).max() < 1e-3
    assert np.array_equal(got.argmax(-1), ref.argmax(-1))
    # against the pure fp32 restatement: report-level bound (bf16 storage), labels still equal here
    assert np.abs(got - ref32).max() < 2e-3
    assert np.array_equal(got.argmax(-1), ref32.argmax(-1))
    assert np.allclose(got.sum(-1), 1.0, atol=1e-5)


def test_predict_batches_and_flatten(setup):
    m = setup["model"]
    full = m.predict(setup["x"], batch_size=4)
    parts = m.predict(setup["x"], batch_size=3)  # ragged last batch
    assert np.array_equal(full, parts)


def test_train_step_loss_grads_and_moving_stats(setup):
    import torch
    import bringup_unet as bu
    from oracle.unet import UNetOracle, filters_for
    m, P, x, y, sw = setup["model"], setup["P"], setup["x"], setup["y"], setup["sw"]
    m.set_keras_weights(P)
    loss_dev = m.forward_backward(x, y, sw)
    torch.cuda.synchronize()
    loss = float(loss_dev.item()) / (CFG["batch"] * CFG["dim"] ** 2)
    enc, bottom, _ = filters_for(4, CFG["cf"])
    force = {}
    for l in range(5):
        for which in ["a1", "a2", "b"] + (["pooled", "u", "bn1", "c2", "c3", "bn2"] if l < 4 else []):
            arr, border, padc = bu.fetch(m, l, which, CFG["batch"], (enc + [bottom])[l])
            assert border == 0 and padc == 0, (which, l)  # zero borders / padded channels invariant
            force["%s_%d" % (which, l)] = arr
    oracle = UNetOracle(CFG["classes"], CFG["channels"], 4, CFG["cf"], params=P)
    loss_ref, grads_ref, stats = oracle.loss_and_grads(x, y, sw, emulate_bf16=True, force=force)
    assert abs(loss - loss_ref) < 1e-5 * max(1.0, abs(loss_ref))
    grads = m.get_flat_grads_as_keras()
    for key, r in grads_ref.items():
        g = grads[key]
        cos = float((g * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30))
        rel = np.abs(g - r).max() / (np.abs(r).max() + 1e-12)
        assert cos > 0.999 and rel < 0.06, (key, cos, rel)
    W2 = m.get_keras_weights()
    for name, (mean, var) in stats.items():
        assert np.abs(W2[name]["moving_mean"] - (0.99 * P[name]["moving_mean"] + 0.01 * mean)).max() < 1e-5
        assert np.abs(W2[name]["moving_variance"] - (0.99 * P[name]["moving_variance"] + 0.01 * var)).max() < 1e-5


def test_adam_matches_keras_rule(setup):
    import torch
    m = setup["model"]
    m.set_keras_weights(setup["P"])
    m.forward_backward(setup["x"], setup["y"], setup["sw"])
    g = m.grads.clone()
    p0 = m.params.clone()
    m.adam_m.zero_()
    m.adam_v.zero_()
    m.optimizer.iterations = 0
    m.optimizer.lr, m.optimizer.epsilon = 1e-3, 1e-8
    m.apply_gradients()
    torch.cuda.synchronize()
    mm, vv = 0.1 * g, 0.001 * g * g
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    expect = p0 - lr_t * mm / (vv.sqrt() + 1e-8)
    assert float((m.params - expect).abs().max()) < 1e-6


def test_training_reduces_loss():
    from multiplanarunet_b200.models import UNet
    rng = np.random.RandomState(3)
    m = UNet(n_classes=3, dim=32, n_channels=2, complexity_factor=0.125, max_batch=4, training=True, seed=1)
    m.optimizer.lr = 1e-3
    x = rng.randn(4, 32, 32, 2).astype(np.float32)
    y = (x[..., 0] > 0).astype(np.uint8) + (x[..., 1] > 1).astype(np.uint8)
    losses = [m.train_on_batch(x, y) for _ in range(30)]
    assert losses[-1] < 0.6 * losses[0], losses[::5]


def test_staged_backward_equals_single_call(setup):
    """mpu_unet_train_forward + mpu_unet_backward_stage(0..2) (the data-parallel overlap entry points)
    produce the gradients of mpu_unet_train_step; the four all-reduce ranges tile the gradient buffer."""
    import ctypes
    import torch
    from multiplanarunet_b200 import _C
    m = setup["model"]
    m.forward_backward(setup["x"], setup["y"], setup["sw"])
    g_ref = m.grads.clone()
    loss_ref = float(m._loss_dev.item())
    B = m._pack(setup["x"])
    y = torch.as_tensor(setup["y"]).cuda()
    sw = torch.as_tensor(setup["sw"]).cuda()
    st = _C.current_stream()
    H, W, _ = m.img_shape
    gscale = 1.0 if m.loss_scale_mode == "sum" else 1.0 / (B * H * W)
    _C.check(_C.lib.mpu_unet_train_forward(m._h, B, _C.ptr(y), _C.ptr(sw), ctypes.c_float(gscale),
                                           _C.ptr(m._loss_dev), _C.ptr(None), st), "train_forward")
    for stage in range(3):
        _C.check(_C.lib.mpu_unet_backward_stage(m._h, B, stage, st), "backward_stage")
    torch.cuda.synchronize()
    assert abs(float(m._loss_dev.item()) - loss_ref) <= 1e-6 * abs(loss_ref)
    # fp32 atomics in the weight-gradient kernels reorder sums between runs
    err = (m.grads - g_ref).abs().max().item()
    assert err <= 1e-4 * g_ref.abs().max().item(), err
    r = (ctypes.c_longlong * 8)()
    _C.check(_C.lib.mpu_unet_grad_ranges(m._h, r), "grad_ranges")
    spans = sorted((r[2 * i], r[2 * i + 1]) for i in range(4))
    assert spans[0][0] == 0 and spans[-1][1] == m.grads.numel()
    assert all(spans[i][1] == spans[i + 1][0] for i in range(3))
    assert _C.lib.mpu_unet_backward_stage(m._h, B, 3, st) != 0


def test_fused_step_with_overlapped_adam_equals_separate_calls():
    """mpu_unet_train_step_adam (Adam of each parameter range on a third stream while the next backward stage runs)
    == forward_backward() + apply_gradients(): same weights, Adam moments and bf16 operand copies after 3 steps, up to
    the run-to-run reordering of the fp32 / fp64 atomics in the gradient kernels."""
    import torch
    from multiplanarunet_b200.models import UNet
    rng = np.random.RandomState(11)
    x = rng.randn(4, 32, 32, 2).astype(np.float32)
    y = (x[..., 0] > 0).astype(np.uint8) + (x[..., 1] > 1).astype(np.uint8)
    kw = dict(n_classes=3, dim=32, n_channels=2, complexity_factor=0.125, max_batch=4, training=True, seed=4)
    for l2 in (None, 1e-4):
        a, b = UNet(l2_reg=l2, **kw), UNet(l2_reg=l2, **kw)
        assert torch.equal(a.params, b.params)
        la = lb = None
        for _ in range(3):
            la = a.train_on_batch(x, y)                 # fused entry
            b.forward_backward(x, y)
            b.apply_gradients()
            H, W, _ = b.img_shape
            lb = float(b._loss_dev.item()) / (4 * H * W) + (l2 * float(b._l2_sumsq[0]) if l2 else 0.0)
        torch.cuda.synchronize()
        assert abs(la - lb) <= 1e-5 * abs(lb), (la, lb)
        # Adam normalises the update, so a parameter whose gradient is pure summation noise may move by up to lr per
        # step in either run; everything else must agree closely
        d = (a.params - b.params).abs()
        assert float(d.max()) <= 3 * 1.01 * a.optimizer.lr, float(d.max())
        assert float((d > 1e-5).float().mean()) < 1e-3
        assert float((a.adam_m - b.adam_m).abs().max()) <= 1e-3 * float(b.adam_m.abs().max())
        # the forward pass of both models agrees (bf16 operand copies and derived up-conv weights were refreshed)
        pa, pb = a.predict_on_batch(x, as_numpy=False), b.predict_on_batch(x, as_numpy=False)
        assert float((pa - pb).abs().max()) < 5e-3


def test_epilogue_reductions_match_separate_passes(monkeypatch):
    """MPU_EPI_RED_LEVEL=0 moves BatchNorm backward's sum g / sum g*y and the conv bias gradients into the epilogue of
    the GEMM that writes the gradient tensor (mtgemm_fwd_kernel<.., RED>); off by default because it measured slower
    (profiles/r02_epilogue_reductions.txt).  Both routes must give the same gradients."""
    import torch
    from multiplanarunet_b200.models import UNet
    rng = np.random.RandomState(21)
    x = rng.randn(4, 64, 64, 1).astype(np.float32)
    y = rng.randint(0, 4, size=(4, 64, 64)).astype(np.uint8)
    kw = dict(n_classes=4, dim=64, n_channels=1, complexity_factor=0.25, max_batch=4, training=True, seed=6)
    a = UNet(**kw)
    monkeypatch.setenv("MPU_EPI_RED_LEVEL", "0")
    b = UNet(**kw)
    monkeypatch.delenv("MPU_EPI_RED_LEVEL")
    la = float(a.forward_backward(x, y).item())
    lb = float(b.forward_backward(x, y).item())
    torch.cuda.synchronize()
    assert abs(la - lb) <= 1e-6 * abs(la)
    ga, gb = a.grads, b.grads
    for info in a._infos:  # per tensor: kernels / gamma at off0, biases / beta at off1
        n0 = info["ksize"] ** 2 * info["co_phys"] * info["k_phys"] if info["kind"] == 0 else info["co_phys"]
        for off, n in ((info["off0"], n0), (info["off1"], info["co_phys"])):
            ra, rb = ga[off:off + n], gb[off:off + n]
            scale = float(ra.abs().max())
            assert float((ra - rb).abs().max()) <= 2e-3 * scale + 1e-6, (info["name"], off)


def test_l2_reg_matches_keras_kernel_regularizer():
    """kernel_regularizer=l2(l2_reg) on every encoder / bottom / up conv kernel, none on the 1x1 head, biases or BN
    (mpunet/models/unet.py:122-189): penalty gradient next to the sum-of-pixels data gradient is (B*H*W) * 2*l2*w,
    and the reported mean loss gains l2 * sum(w^2)."""
    import torch
    from multiplanarunet_b200.models import UNet
    rng = np.random.RandomState(5)
    l2 = 1e-3
    m = UNet(n_classes=3, dim=32, n_channels=1, complexity_factor=0.125, max_batch=2, training=True, seed=2, l2_reg=l2)
    x = rng.randn(2, 32, 32, 1).astype(np.float32)
    y = rng.randint(0, 3, size=(2, 32, 32)).astype(np.uint8)
    m.forward_backward(x, y)
    torch.cuda.synchronize()
    m.grads.zero_()  # (adding onto the data gradient would only test fp32 rounding of g + penalty)
    m._l2_begin()
    m._l2_penalty(0, m.grads.numel())
    torch.cuda.synchronize()
    diff = m.grads.clone()
    expect = torch.zeros_like(diff)
    sumsq = 0.0
    for info in m._infos:
        if info["kind"] != 0 or info["name"] == "conv2d":
            continue
        n = info["ksize"] ** 2 * info["co_phys"] * info["k_phys"]
        w = m.params[info["off0"]:info["off0"] + n]
        expect[info["off0"]:info["off0"] + n] = 2 * l2 * (2 * 32 * 32) * w
        sumsq += float((w.double() ** 2).sum())
    assert sumsq > 0
    assert float((diff - expect).abs().max()) <= 1e-6 * float(expect.abs().max())
    assert abs(float(m._l2_sumsq[0]) - sumsq) <= 1e-6 * sumsq
    # the public step reports data loss + penalty and still trains
    m2 = UNet(n_classes=3, dim=32, n_channels=1, complexity_factor=0.125, max_batch=2, training=True, seed=2)
    la, lb = m.train_on_batch(x, y), m2.train_on_batch(x, y)
    assert abs((la - lb) - l2 * sumsq) <= 1e-5 * max(1.0, abs(la))
