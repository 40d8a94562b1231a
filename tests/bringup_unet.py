"""GPU bring-up harness for the U-Net engine vs the CPU oracle (run on the B200 box via gpurun).
  python tests/bringup_unet.py [--dim 64 --batch 4 --cf 0.125 --classes 3 --channels 1]
Prints per-activation and per-gradient error tables so a wrong kernel can be localised in one run.
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WHICH = {"a1": 0, "a2": 1, "b": 2, "pooled": 3, "u": 4, "bn1": 5, "c2": 6, "c3": 7, "bn2": 8}


def fetch(model, level, which, B, C_logical):
    import torch
    from multiplanarunet_b200._C import lib, check
    ptr, rows, C = ctypes.c_void_p(), ctypes.c_longlong(), ctypes.c_int()
    check(lib.mpu_unet_debug_buffer(model._h, level, WHICH[which], ctypes.byref(ptr), ctypes.byref(rows),
                                    ctypes.byref(C)))
    H = model.img_shape[0] >> level
    W = model.img_shape[1] >> level
    if which == "pooled":
        H, W = H // 2, W // 2
    n = B * (H + 2) * (W + 2) * C.value
    torch.cuda.synchronize()
    # the buffer lives inside the torch-owned workspace: view it through its byte offset
    off = ptr.value - model.workspace.data_ptr()
    src = model.workspace[off:off + 2 * n].view(torch.bfloat16)
    arr = src.float().cpu().numpy().reshape(B, H + 2, W + 2, C.value)
    bm = np.ones(arr.shape[:3], dtype=bool)
    bm[:, 1:-1, 1:-1] = False
    border = float(np.abs(arr[bm]).sum())
    padc = np.abs(arr[..., C_logical:]).sum()
    return arr[:, 1:-1, 1:-1, :C_logical], border, padc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--cf", type=float, default=0.125)
    ap.add_argument("--classes", type=int, default=3)
    ap.add_argument("--channels", type=int, default=1)
    ap.add_argument("--skip-train", action="store_true")
    ap.add_argument("--skip-acts", action="store_true")
    args = ap.parse_args()
    import torch
    from multiplanarunet_b200.models import UNet
    from oracle.unet import UNetOracle, init_params, filters_for

    rng = np.random.RandomState(0)
    P = init_params(args.classes, args.channels, 4, args.cf, seed=1, randomize_bn=True)
    for n, d in P.items():  # non-trivial biases
        if "bias" in d:
            d["bias"] = (0.05 * rng.randn(*d["bias"].shape)).astype(np.float32)
    x = rng.randn(args.batch, args.dim, args.dim, args.channels).astype(np.float32)
    y = rng.randint(0, args.classes, size=(args.batch, args.dim, args.dim)).astype(np.uint8)
    sw = rng.uniform(0.5, 1.5, size=args.batch).astype(np.float32)

    model = UNet(n_classes=args.classes, dim=args.dim, n_channels=args.channels, complexity_factor=args.cf,
                 max_batch=args.batch, training=True)
    model.set_keras_weights(P)
    oracle = UNetOracle(args.classes, args.channels, 4, args.cf, params=P)
    print("count_params: model %d oracle-trainable+bnstats" % model.count_params())

    # round trip of the weight layout
    P2 = model.get_keras_weights()
    rt = max(np.abs(P[n][k] - P2[n][k]).max() for n in P for k in P[n])
    print("weight round-trip max diff:", rt)

    enc, bottom, up = filters_for(4, args.cf)
    chans = enc + [bottom]

    def compare_acts(tag, training):
        cap = {}
        ref = oracle.predict(x, emulate_bf16=True, training=training, capture=cap)
        got = model.predict_on_batch(x, bn_training=training)
        ref32 = oracle.predict(x, emulate_bf16=False, training=training)
        print("[%s] probs: max|gpu-oracle_bf16emu|=%.3g  max|gpu-oracle_fp32|=%.3g  max|emu-fp32|=%.3g "
              "argmax mismatch vs emu=%.4g vs fp32=%.4g" %
              (tag, np.abs(got - ref).max(), np.abs(got - ref32).max(), np.abs(ref - ref32).max(),
               (got.argmax(-1) != ref.argmax(-1)).mean(), (got.argmax(-1) != ref32.argmax(-1)).mean()))
        names = []
        for l in range(5):
            names += [("a1", l), ("a2", l), ("b", l)] + ([("pooled", l)] if l < 4 else [])
        for l in (3, 2, 1, 0):
            names += [("u", l), ("bn1", l), ("c2", l), ("c3", l), ("bn2", l)]
        for which, l in names:
            key = "%s_%d" % (which, l)
            r = cap[key]
            g, border, padc = fetch(model, l, which, args.batch, r.shape[-1])
            err = np.abs(g - r)
            tol = 0.02 + 0.02 * np.abs(r)
            print("   %-10s max_err=%.4g mean_err=%.3g frac_bad=%.4g ref_absmean=%.3g border=%.3g padch=%.3g" %
                  (key, err.max(), err.mean(), (err > tol).mean(), np.abs(r).mean(), border, padc))
        return got, ref

    if not args.skip_acts:
        compare_acts("inference", False)
        compare_acts("bn-training-forward", True)

    if args.skip_train:
        return
    # ---- train step: loss + gradients
    model.set_keras_weights(P)  # reset moving stats changed by the bn-training forward
    loss_dev = model.forward_backward(x, y, sw)
    torch.cuda.synchronize()
    loss_gpu = float(loss_dev.item()) / (args.batch * args.dim * args.dim)
    oracle = UNetOracle(args.classes, args.channels, 4, args.cf, params=P)
    force = {}
    for l in range(5):
        for which in ["a1", "a2", "b"] + (["pooled", "u", "bn1", "c2", "c3", "bn2"] if l < 4 else []):
            force["%s_%d" % (which, l)] = fetch(model, l, which, args.batch,
                                                 (enc + [bottom])[l])[0]
    loss_ref, grads_ref, stats = oracle.loss_and_grads(x, y, sw, emulate_bf16=True, force=force)
    loss_32, grads_32, _ = oracle.loss_and_grads(x, y, sw, emulate_bf16=False)
    print("loss: gpu=%.6f oracle_emu=%.6f oracle_fp32=%.6f" % (loss_gpu, loss_ref, loss_32))
    grads = model.get_flat_grads_as_keras()
    worst = 0
    for key in grads_ref:
        g, r, r32 = grads[key], grads_ref[key], grads_32[key]
        denom = np.abs(r).max() + 1e-12
        rel = np.abs(g - r).max() / denom
        rel32 = np.abs(g - r32).max() / (np.abs(r32).max() + 1e-12)
        emu32 = np.abs(r - r32).max() / (np.abs(r32).max() + 1e-12)
        cos = float((g * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30))
        worst = max(worst, rel)
        print("   grad %-28s rel_maxerr(vs emu)=%.3g (vs fp32)=%.3g [emu-vs-fp32 %.3g] cos=%.6f |ref|max=%.3g" %
              ("%s/%s" % key, rel, rel32, emu32, cos, denom))
    print("worst rel grad err vs emu oracle: %.4g" % worst)
    # moving stats after one training forward
    W2 = model.get_keras_weights()
    for name, (m, v) in list(stats.items())[:3] + list(stats.items())[-2:]:
        mm = 0.99 * P[name]["moving_mean"] + 0.01 * m
        mv = 0.99 * P[name]["moving_variance"] + 0.01 * v
        print("   moving stats %-22s mean err %.3g var err %.3g" %
              (name, np.abs(W2[name]["moving_mean"] - mm).max(), np.abs(W2[name]["moving_variance"] - mv).max()))
    # one Adam step moves the weights
    model.optimizer.lr = 1e-3
    before = model.params.clone()
    model.apply_gradients()
    torch.cuda.synchronize()
    delta = (model.params - before).abs()
    print("adam: max |delta| = %.4g (lr 1e-3), nonzero frac = %.4g" % (delta.max().item(), (delta > 0).float().mean().item()))


if __name__ == "__main__":
    main()
