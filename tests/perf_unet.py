"""Quick device-timed profile of the U-Net train step at the benchmark size (run via gpurun).
  python tests/perf_unet.py [--dim 256 --batch 32 --cf 2 --steps 5]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FWD_GFLOP = {2.0: 217.59, 1.0: 109.10}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--cf", type=float, default=2.0)
    ap.add_argument("--classes", type=int, default=5)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--ncu", action="store_true", help="profile exactly one step (use with ncu --profile-from-start off)")
    args = ap.parse_args()
    import torch
    from multiplanarunet_b200.models import UNet
    model = UNet(n_classes=args.classes, dim=args.dim, n_channels=1, complexity_factor=args.cf,
                 max_batch=args.batch, training=True, seed=0)
    print("workspace GB: %.2f  params: %d" % (model.workspace.numel() / 1e9, model.count_params()))
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(args.batch, args.dim, args.dim, 1, device="cuda", generator=g)
    y = torch.randint(0, args.classes, (args.batch, args.dim, args.dim), device="cuda", generator=g,
                      dtype=torch.uint8)

    def timed(fn, n):
        evs = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    def step():
        model.train_on_batch_async(x, y)  # the product step: forward + backward + Adam per range

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if args.ncu:
        torch.cuda.profiler.start()
        step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    t_step = timed(step, args.steps)
    t_fb = timed(lambda: model.forward_backward(x, y), args.steps)
    t_fwd = timed(lambda: model.predict_on_batch(x, bn_training=True, as_numpy=False), args.steps)
    t_adam = timed(model.apply_gradients, args.steps)
    fl = FWD_GFLOP.get(args.cf, 0) * (args.dim / 256.0) ** 2 * args.batch
    ms = float(np.median(t_step))
    print("train step ms: median %.2f  all %s" % (ms, ["%.2f" % t for t in t_step]))
    print("  fwd+bwd ms %.2f | fwd(train-BN)+head ms %.2f | adam+weight-sync ms %.2f" %
          (np.median(t_fb), np.median(t_fwd), np.median(t_adam)))
    print("  slices/s %.1f   conv TFLOP/s (3x fwd algorithmic) %.1f" %
          (args.batch / ms * 1e3, 3 * fl / ms))
    print("  fwd-only TFLOP/s %.1f" % (fl / float(np.median(t_fwd))))
    print("loss", float(model._loss_dev.item()) / (args.batch * args.dim * args.dim))


if __name__ == "__main__":
    main()
