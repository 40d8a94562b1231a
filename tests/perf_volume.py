"""Device timing of the volume-side hot kernels at BASELINE config 3 / 5 sizes (run on the B200 box):
  * sample_planes_kernel : one view's inference stack, 276 planes of 256x256 from a resident 256^3 volume
  * map_fuse_kernel      : 6 views x 276 planes x 256^2 x 5 classes -> 256^3 labels
  * fusion_grad_kernel   : one epoch pass over N = 256^3 points x 6 views x 5 classes (121 B/point)
Prints ms and GB/s on the ALGORITHMIC bytes of SURVEY.md 8(d); `--once` runs each kernel a single time after
one warm-up (the shape used under `ncu --set full`).
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, iters, flush=None):
    import torch
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)  # > L2: evicts the previous iteration's lines
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--classes", type=int, default=5)
    ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from multiplanarunet_b200 import _C
    from multiplanarunet_b200._C import lib, check
    from multiplanarunet_b200.interpolation import (ViewInterpolator, plane_basis, view_offsets,
                                                    sample_random_views_with_angle_restriction)
    from multiplanarunet_b200.utils.fusion.fuse_and_predict import _map_fuse
    dim, C, V = args.dim, args.classes, args.views
    iters = 1 if args.once else args.iters
    dev = torch.device("cuda")
    flush = None if args.once else torch.zeros(64 * 1024 * 1024, device=dev)  # 256 MB > 126 MB L2
    res = {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)

    # ---- sampler
    g = torch.Generator(device=dev).manual_seed(0)
    vol = torch.randn(dim, dim, dim, 1, device=dev, generator=g)
    lab = torch.randint(0, C, (dim, dim, dim), device=dev, dtype=torch.uint8)
    vi = ViewInterpolator(vol, lab, np.eye(4), bg_value=[-1.0], bg_class=0)
    np.random.seed(0)
    views = sample_random_views_with_angle_restriction(V, 60)
    offs = view_offsets(dim, float(dim), "same+20")
    n = len(offs)
    pad = torch.zeros(n * (dim + 2) * (dim + 2), 8, dtype=torch.bfloat16, device=dev)
    basis = plane_basis(views[0], 0.)

    def run_sampler():
        vi.sample_planes(basis, offs, dim, float(dim), center=[0.1], scale=[1.3], out_padded=pad, cpad=8,
                         want_f32=False, want_labels=False)
    ms, all_ms = timeit(run_sampler, iters, flush)
    npts = n * dim * dim
    alg = dim ** 3 * 4 + npts * 2       # read the volume once + write bf16 slices (SURVEY 8d lower bound)
    res["sample_planes"] = {"ms": ms, "points": npts, "gpoints_per_s": npts / ms / 1e6,
                            "algorithmic_gbs": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / hbm,
                            "written_gbs_incl_padding": npts * 16 / ms / 1e6}
    print("sample_planes: %.3f ms for %d planes (%.2f Gpoint/s; algorithmic %.1f GB/s)" % (
        ms, n, npts / ms / 1e6, alg / ms / 1e6))

    # ---- map + fuse
    pred_dtype = torch.float32
    preds = []
    for v in range(V):
        p = torch.rand(n, dim, dim, C, device=dev, generator=g)
        preds.append((p / p.sum(-1, keepdim=True)).to(pred_dtype))
    gax = np.linspace(-(dim // 2), dim // 2, dim)
    grids = [(gax, gax, offs)] * V
    ibs = [np.linalg.inv(plane_basis(v, 0.)) for v in views]
    W = np.random.RandomState(0).uniform(0.5, 1.5, (V, C)).astype(np.float32)
    b = np.zeros(C, np.float32)

    def run_fuse():
        _map_fuse(preds, grids, ibs, (dim, dim, dim), np.eye(3), W, b)
    ms, all_ms = timeit(run_fuse, iters, flush)
    alg = dim ** 3 * (V * C * 4 + 1)
    res["map_fuse"] = {"ms": ms, "algorithmic_gbs": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / hbm,
                       "bytes_per_voxel": V * C * 4 + 1, "pred_dtype": str(pred_dtype)}
    print("map_fuse: %.3f ms (%.1f GB/s on %d B/voxel = %.3f of measured HBM)" % (
        ms, alg / ms / 1e6, V * C * 4 + 1, alg / ms / 1e6 / hbm))
    del preds

    # ---- fusion training pass
    N = dim ** 3
    X = torch.rand(N, V, C, device=dev, generator=g)
    y = torch.randint(0, C, (N,), device=dev, dtype=torch.uint8)
    Wd = torch.as_tensor(W).to(dev)
    bd = torch.as_tensor(b).to(dev)
    acc = torch.zeros(V * C + C + 1, dtype=torch.float64, device=dev)

    def run_fgrad():
        check(lib.mpu_fusion_grad(_C.ptr(X), _C.ptr(y), ctypes.c_longlong(N), V, C, _C.ptr(Wd), _C.ptr(bd),
                                  _C.ptr(acc), _C.current_stream()), "mpu_fusion_grad")
    ms, all_ms = timeit(run_fgrad, iters, flush)
    alg = N * (V * C * 4 + 1)
    res["fusion_grad"] = {"ms": ms, "algorithmic_gbs": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / hbm,
                          "points": N}
    print("fusion_grad: %.3f ms per pass over %d points (%.1f GB/s = %.3f of measured HBM)" % (
        ms, N, alg / ms / 1e6, alg / ms / 1e6 / hbm))
    res["hbm_peak_gbs"] = hbm
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
