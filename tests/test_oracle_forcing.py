"""CPU self-consistency of the oracle's teacher-forcing hooks (oracle/unet.py): forcing the oracle with its OWN
activations and its OWN activation gradients must reproduce the unforced run exactly, every layer-local residual
must be zero, and the taps must see the gradient at every point the device materialises."""
import numpy as np

from oracle.unet import UNetOracle, init_params


def test_forcing_with_own_tensors_is_identity():
    rng = np.random.RandomState(0)
    P = init_params(3, 1, 4, 0.125, seed=1, randomize_bn=True)
    x = rng.randn(2, 32, 32, 1).astype(np.float32)
    y = rng.randint(0, 3, size=(2, 32, 32))
    sw = np.array([0.7, 1.2], np.float32)
    cap, seen0 = {}, {}
    loss0, g0, _ = UNetOracle(3, 1, 4, 0.125, params=P).loss_and_grads(x, y, sw, emulate_bf16=True, grad_seen=seen0,
                                                                      capture=cap)
    names = set("%s_%d" % (w, l) for l in range(4) for w in ("a1", "a2", "b", "pooled", "u", "bn1", "c2", "c3", "bn2"))
    names |= {"a1_4", "a2_4", "b_4"}
    assert set(cap) == names
    assert set(seen0) == names | {"skip_%d" % l for l in range(4)}
    computed, seen1 = {}, {}
    fg = {k: v for k, v in seen0.items() if not (k.startswith("b_") and k != "b_4")}
    loss1, g1, _ = UNetOracle(3, 1, 4, 0.125, params=P).loss_and_grads(
        x, y, sw, emulate_bf16=True, force=cap, computed=computed, grad_seen=seen1, force_grad=fg)
    assert loss1 == loss0
    for k in cap:
        assert np.array_equal(computed[k], cap[k]), k
    for k in seen0:
        assert np.array_equal(seen1[k], seen0[k]), k
    for k in g0:
        assert np.array_equal(g0[k], g1[k]), k
    # gradients at activations the device stores in bf16 are bf16 values in the emulating oracle
    import torch
    for k in ("a1_2", "pooled_1", "bn1_0", "skip_3", "u_1", "bn2_2"):
        t = torch.as_tensor(seen0[k])
        assert torch.equal(t.to(torch.bfloat16).float(), t), k
