#!/usr/bin/env python
"""Benchmark of the mpunet hot path on B200 (driver contract: python bench.py --gpus N --steps K --warmup W).

Workload (BASELINE.json configs[1]): 6-view 2D U-Net train step on 256x256 oblique slices, batch 32
per GPU, bf16 GEMMs / fp32 accumulate, complexity_factor 2 (the reference's default YAML), 5 classes,
slices sampled on the device from a resident synthetic 256^3 volume (one volume per GPU).
A step = sample 32 oblique planes (trilinear image + nearest labels + RobustScaler) straight into the
U-Net input tensor -> forward -> sparse-CE -> backward -> (NCCL all-reduce of the fp32 gradients when
N > 1) -> Adam.  Metric: slices/s, whole job.

  value : inputs resident in HBM (volume on the device, planes drawn on the host beforehand)
  e2e   : the reference-facing call model.train_on_batch(x, y, w) with HOST (pinned) numpy batches:
          H2D copies of x / y / w and the D2H loss read are inside the timed region
  --impl reference : the CPU restatement of the reference's Keras train step (oracle/unet.py; TensorFlow
          cannot be installed here) on the host cores, bounded to 1 slice per step
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GFLOP_PER_SLICE = {2.0: 217.59, 1.0: 109.10}  # conv MACs*2, unpadded channels, 256x256, 5 classes
REF_BATCH = 1


def synthetic_volume(dim, n_classes, seed, device):
    """Smooth random field + ellipsoid labels (a la mpunet/bin/toy_data.py), generated with torch."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    low = torch.randn(1, 1, dim // 8, dim // 8, dim // 8, device=device, generator=g)
    vol = torch.nn.functional.interpolate(low, size=(dim, dim, dim), mode="trilinear", align_corners=False)[0, 0]
    vol = vol * 100.0 + 10.0 * torch.randn(dim, dim, dim, device=device, generator=g)
    ax = torch.arange(dim, device=device, dtype=torch.float32)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    labels = torch.zeros(dim, dim, dim, dtype=torch.uint8, device=device)
    rng = np.random.RandomState(seed)
    for c in range(1, n_classes):
        ctr = rng.uniform(0.25 * dim, 0.75 * dim, 3)
        rad = rng.uniform(0.12 * dim, 0.25 * dim, 3)
        inside = ((X - ctr[0]) / rad[0]) ** 2 + ((Y - ctr[1]) / rad[1]) ** 2 + ((Z - ctr[2]) / rad[2]) ** 2 < 1
        labels[inside] = c
        vol = vol + inside.float() * (40.0 * c)
    return vol.unsqueeze(-1).contiguous(), labels.contiguous()


class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, reasons, smax = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [s for s in sm if smax and s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step_factory(cf, dim, n_classes):
    """The reference's train step restated on the CPU (oracle/unet.py): fwd + sparse CE + bwd + Adam."""
    import torch
    from oracle.unet import UNetOracle
    # torch's default intra-op pool = all physical cores the process may use (forcing every logical CPU
    # oversubscribed the 128-thread GPU host: 57 s per 2 slices)
    oracle = UNetOracle(n_classes, 1, 4, cf, seed=0)
    rng = np.random.RandomState(0)
    x = rng.randn(REF_BATCH, dim, dim, 1).astype(np.float32)
    y = rng.randint(0, n_classes, size=(REF_BATCH, dim, dim))
    state = {"t": 0, "m": {}, "v": {}}

    def step():
        _, grads, _ = oracle.loss_and_grads(x, y)
        state["t"] += 1
        t = state["t"]
        lr_t = 5e-5 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        for (name, key), g in grads.items():
            gt = torch.from_numpy(g)
            m = state["m"].setdefault((name, key), torch.zeros_like(gt))
            v = state["v"].setdefault((name, key), torch.zeros_like(gt))
            m.mul_(0.9).add_(gt, alpha=0.1)
            v.mul_(0.999).addcmul_(gt, gt, value=0.001)
            oracle.P[name][key].sub_(lr_t * m / (v.sqrt() + 1e-8))
    return step, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, cores = cpu_reference_step_factory(args.cf, args.dim, args.classes)
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = REF_BATCH * args.steps / dt
    sample = "%d train steps (fwd + sparse-CE + bwd + Adam) of the torch-CPU fp32 restatement on %d slices each" % (
        args.steps, REF_BATCH)
    print(json.dumps({
        "impl": "reference", "metric": "slices_per_sec", "value": val, "unit": "slices/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, REF_BATCH, "cpu"),
        "cpu_baseline": {"value": val, "unit": "slices/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, batch, where):
    return {"workload": "6-view 2D U-Net train step on %dx%d oblique slices of a synthetic %d^3 volume" % (
        args.dim, args.dim, args.dim), "slices_per_step_per_gpu": batch, "complexity_factor": args.cf,
        "n_classes": args.classes, "n_views": 6, "optimizer": "Adam", "where": where,
        "l2_policy": "per-step working set (>10 GB of activations) far exceeds the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cf", type=float, default=2.0)
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--classes", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a short collective timeout: a rank-asymmetric bug must abort in minutes, not hang the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    from multiplanarunet_b200 import _C
    from multiplanarunet_b200._C import lib, check
    from multiplanarunet_b200.interpolation import plane_basis, sample_random_views_with_angle_restriction
    from multiplanarunet_b200.models import UNet
    from multiplanarunet_b200.sequences import SyntheticImage
    lib.mpu_launch_count.restype = ctypes.c_longlong

    B, dim, K, W = args.batch, args.dim, args.steps, args.warmup
    # ---- data: one resident synthetic volume per GPU, 6 fixed views shared by all ranks
    vol, labels = synthetic_volume(dim, args.classes, 1000 + rank, dev)
    image = SyntheticImage(vol.cpu().numpy(), labels.cpu().numpy(), np.eye(4), device=dev)
    del vol, labels
    np.random.seed(0)
    views = sample_random_views_with_angle_restriction(6, 60)
    model = UNet(n_classes=args.classes, dim=dim, n_channels=1, complexity_factor=args.cf, max_batch=B,
                 training=True, seed=0, device=dev)
    model.optimizer.lr, model.optimizer.epsilon = 5e-5, 1e-8
    span = float(dim)
    # planes for every step are drawn on the host before the clock starts (view, offset ~ U(-span//2,
    # span//2), normal noise sd 0.1 as in sequences/isotrophic_live_view_sequence_2d.py:119-141)
    rng = np.random.RandomState(100 + rank)
    n_steps_total = W + K + 2
    planes = np.empty((n_steps_total, B, 10))
    for s in range(n_steps_total):
        for b in range(B):
            planes[s, b, :9] = plane_basis(views[rng.randint(0, 6)], rng.normal(scale=0.1, size=3)).ravel()
            planes[s, b, 9] = rng.uniform(-(span // 2), span // 2)
    planes_d = torch.from_numpy(planes).to(dev)
    in_ptr, cpad, rows = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_longlong()
    check(lib.mpu_unet_input_buffer(model._h, ctypes.byref(in_ptr), ctypes.byref(cpad), ctypes.byref(rows)))
    y_dev = torch.empty(B, dim, dim, dtype=torch.uint8, device=dev)
    interp = image.interpolator
    dims = _C.int_array(interp.im_shape[:3])
    bg = _C.float_array(interp.bg_value)
    cen, scl = _C.double_array(image.scaler_center), _C.double_array(image.scaler_scale)

    def sample_into_unet(s):
        check(lib.mpu_sample_planes(_C.ptr(interp.vol), _C.ptr(interp.labels), dims, 1, _C.ptr(interp._g[0]),
                                    _C.ptr(interp._g[1]), _C.ptr(interp._g[2]), interp._inv_step, None,
                                    ctypes.c_void_p(planes_d[s].data_ptr()), B, dim, ctypes.c_double(span), bg, 0,
                                    cen, scl, None, in_ptr, cpad.value, _C.ptr(y_dev), _C.current_stream()),
              "mpu_sample_planes")

    def step_value(s):
        sample_into_unet(s)
        # SUM all-reduce of the gradients (MirroredStrategy's aggregation) overlapped with backward
        model.forward_backward_overlapped(None, y_dev, None, input_packed=True, batch=B)
        model.apply_gradients()

    def timed(fn, nwarm, nsteps, first):
        for i in range(nwarm):
            fn(first + i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = lib.mpu_launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(nsteps):
            fn(first + nwarm + i)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        launches = lib.mpu_launch_count() - l0
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(t.item()), launches

    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
        time.sleep(0.3)
    ms_value, launches = timed(step_value, W, K, 0)
    clock_info = clocks.stop() if clocks else None

    # ---- e2e: host (pinned) numpy batches through the Keras-like call
    x_f32 = torch.empty(B, dim, dim, 1, dtype=torch.float32, device=dev)
    check(lib.mpu_sample_planes(_C.ptr(interp.vol), _C.ptr(interp.labels), dims, 1, _C.ptr(interp._g[0]),
                                _C.ptr(interp._g[1]), _C.ptr(interp._g[2]), interp._inv_step, None,
                                ctypes.c_void_p(planes_d[W + K].data_ptr()), B, dim, ctypes.c_double(span), bg, 0,
                                cen, scl, _C.ptr(x_f32), None, 0, _C.ptr(y_dev), _C.current_stream()))
    x_host = x_f32.cpu().pin_memory()
    y_host = y_dev.cpu().pin_memory()
    w_host = torch.ones(B, dtype=torch.float32).pin_memory()
    losses = []

    def step_e2e(_):
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        wd = w_host.to(dev, non_blocking=True)
        loss = model.forward_backward_overlapped(xd, yd, wd)
        model.apply_gradients()
        losses.append(float(loss.item()) / (B * dim * dim))  # D2H read of the step's loss

    ms_e2e, _ = timed(step_e2e, 2, K, 0)
    h2d = x_host.numel() * 4 + y_host.numel() + w_host.numel() * 4
    d2h = 8

    # ---- roofline of the dominant kernels (tensor-core GEMMs), timed live with CUDA events
    # (every rank runs the extra step - it contains the gradient all-reduce - only rank 0 reads the timer)
    roofline = None
    check(lib.mpu_profile_gemm(1))
    step_value(W + K + 1)
    gemm_ms, n_l = ctypes.c_double(), ctypes.c_int()
    check(lib.mpu_profile_gemm_read(ctypes.byref(gemm_ms), ctypes.byref(n_l)))
    check(lib.mpu_profile_gemm(0))
    if rank == 0:
        # algorithmic FLOPs executed by the tensor-core GEMM kernels: 3x forward minus the first conv
        # (CUDA-core kernels, no input gradient) and the 1x1 head (fused softmax/CE kernel)
        non_gemm = {2.0: 0.106 + 0.059, 1.0: 0.075 + 0.042}.get(args.cf, 0.0)
        flops = 3.0 * (FWD_GFLOP_PER_SLICE.get(args.cf, 0.0) - non_gemm) * 1e9 * (dim / 256.0) ** 2 * B
        # DRAM bytes per GEMM launch from the committed ncu pass over one train step (profiles/)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")))
            if args.cf == 2.0 and dim == 256 and B == 32:
                traffic = tr["gemm_dram_bytes_per_launch"]
        except Exception:
            pass
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained")
        src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long step)"
        if not peak:
            peak, src = 1400.0, "fallback (B200_PROFILING.md sustained figure)"
        ach = flops / (gemm_ms.value * 1e-3) / 1e12 if gemm_ms.value > 0 else None
        roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": (ach / peak) if ach else None, "traffic": traffic,
                    "traffic_note": "mean dram__bytes_read+write per GEMM launch, ncu pass over one train step "
                                    "(profiles/r01_gemm_traffic.json)" if traffic else None,
                    "kernels": "mtgemm_fwd_kernel + mtgemm_wgrad_kernel (%d launches/step, %.2f ms/step)" % (
                        n_l.value, gemm_ms.value),
                    "algorithmic_flops_per_step": flops, "peak_source": src}

    # ---- BASELINE.json's secondary figures (config 3): 6-view predict + fusion on the resident volume
    extras = None
    if rank == 0 and world == 1:
        from multiplanarunet_b200.sequences import IsotrophicLiveViewSequence2D
        from multiplanarunet_b200.utils.fusion.fuse_and_predict import _map_fuse, predict_stack_device
        seq = IsotrophicLiveViewSequence2D([image], views=views, sample_dim=dim, real_space_span=span,
                                           n_classes=args.classes, is_validation=True)
        Wf = np.random.RandomState(0).uniform(0.5, 1.5, (6, args.classes)).astype(np.float32)
        bf = np.zeros(args.classes, np.float32)

        def predict_volume():
            preds, grids, ibs = [], [], []
            for v in views:
                p_, g_, ib_ = predict_stack_device(model, seq, image, v, "same+20", B)
                preds.append(p_)
                grids.append(g_)
                ibs.append(ib_)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            lab, _, _ = _map_fuse(preds, grids, ibs, image.shape[:3], image.affine[:3, :3], Wf, bf)
            ev[1].record()
            return lab, ev
        predict_volume()  # warm-up
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lab, ev = predict_volume()
        b.record()
        torch.cuda.synchronize()
        ms_vol, ms_fuse = a.elapsed_time(b), ev[0].elapsed_time(ev[1])
        nvox = float(dim) ** 3
        fuse_bytes = nvox * (6 * args.classes * 4 + 1)  # reference-shaped traffic: read V*C fp32, write 1 label
        hbm = peaks.get("hbm_gbs")
        extras = {"predict_volumes_per_sec": 1e3 / ms_vol, "predict_ms_per_volume": ms_vol,
                  "predict_config": "6 views x %d planes of %dx%d, U-Net inference (moving-stat BN) + map + fuse + argmax, "
                                    "one %d^3 volume resident in HBM" % (dim + 20, dim, dim, dim),
                  "fusion_ms": ms_fuse, "fusion_algorithmic_gbs": fuse_bytes / ms_fuse / 1e6,
                  "fusion_frac_of_measured_hbm": (fuse_bytes / ms_fuse / 1e6 / hbm) if hbm else None,
                  "foreground_fraction": float((lab > 0).float().mean().item())}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        step, cores = cpu_reference_step_factory(args.cf, dim, args.classes)
        step()  # warm-up: oneDNN primitive creation, allocator
        n_cpu_steps = 5  # bounded sample: about 10 s of host work at 0.5 slices/s
        t0 = time.perf_counter()
        for _ in range(n_cpu_steps):
            step()
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": REF_BATCH * n_cpu_steps / dt, "unit": "slices/s", "cores": cores, "kind": "port",
                        "sample": "%d train steps (fwd + sparse-CE + bwd + Adam) of oracle/unet.py (torch-CPU fp32) "
                                  "on %d slice(s) each after one warm-up step, %.1f s"
                                  % (n_cpu_steps, REF_BATCH, dt)}

    if rank == 0:
        out = {
            "metric": "slices_per_sec", "value": world * B * K / (ms_value * 1e-3), "unit": "slices/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_value / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, B, "B200"),
            "clocks": clock_info,
            "e2e": {"value": world * B * K / (ms_e2e * 1e-3), "unit": "slices/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "last_loss": losses[-1] if losses else None},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "extras": extras,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
