#!/usr/bin/env python
"""Benchmark of the mpunet hot path on B200 (driver contract: python bench.py --gpus N --steps K --warmup W).

Default workload (BASELINE.json configs[1], `--workload train`): 6-view 2D U-Net train step on 256x256 oblique
slices, batch 32 per GPU, bf16 GEMM operands / fp32 accumulate, complexity_factor 2 (the reference's default YAML),
5 classes, slices sampled on the device from a resident synthetic 256^3 volume (one volume per GPU).
A step = sample 32 oblique planes (trilinear image + nearest labels + RobustScaler) straight into the U-Net input
tensor -> forward -> sparse-CE -> backward -> (NCCL all-reduce of the fp32 gradients when N > 1) -> Adam.
Metric: slices/s, whole job.

  value : inputs resident in HBM (volume on the device, planes drawn on the host before the clock starts)
  e2e   : the reference-facing call `model.train_on_batch(x, y, w)` with HOST (pinned) batches: the H2D copies of
          x / y / w and the D2H read of the loss are inside the timed region
  fit   : (extras) `UNet.fit` fed by the device rejection sampler (foreground rules, is_valid_im, Elastic2D p=1/3)
  --impl reference : the CPU restatement of the reference's Keras train step (oracle/unet.py; TensorFlow cannot be
          installed here) on ALL host cores of the box, 4 slices per step (BASELINE.md 4.3 allows a reduced batch)

Other configurations of BASELINE.json, reported inside the default line at N=1 and selectable as the primary
workload with `--workload`:
  predict      (configs[2]) one 256^3 volume: host volume -> 6 views x 276 planes -> U-Net -> map + fuse -> host
               uint8 label map (H2D 67 MB and D2H 16.8 MB inside the clock), beside the CPU pipeline restatement
               (oracle/predict_pipeline.py) timed on a bounded sample of the same volume
  train_fusion (configs[4]) epochs of the 35-parameter fusion layer over 256^3 x 6 x 5 mapped softmax points per GPU
               (batch 2^17 per rank, shuffled every epoch, 36-double all-reduce per batch at N > 1): GB/s on the
               121 B/point the reference streams
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
REF_BATCH = 4      # slices per CPU reference step (the GPU arm runs 32; slices/s is batch-independent on the CPU)
FUSION_BYTES_PER_POINT = lambda V, C: V * C * 4 + 1  # noqa: E731  (X fp32 + y)


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


# ------------------------------------------------------------------------------------------------ CPU arms
def host_cores():
    """(threads this process may use, logical CPUs, physical cores or None).  torchrun exports OMP_NUM_THREADS=1:
    the CPU arms set their thread count explicitly from the affinity mask instead."""
    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    physical = None
    try:
        import psutil
        physical = psutil.cpu_count(logical=False)
    except Exception:
        pass
    return usable, os.cpu_count(), physical


def use_all_host_threads():
    import torch
    usable, logical, physical = host_cores()
    # every CPU the process may run on; on an unrestricted SMT host one thread per physical core (SMT siblings
    # oversubscribe oneDNN's GEMMs: 64 threads on a 128-thread host were 3x slower than 16 in round 1)
    n = usable
    if physical and logical and usable == logical and physical < logical:
        n = physical
    torch.set_num_threads(max(1, n))
    return n, usable, logical, physical


def cpu_reference_step_factory(cf, dim, n_classes, batch=REF_BATCH):
    """The reference's train step restated on the CPU (oracle/unet.py): fwd + sparse CE + bwd + Adam."""
    import torch
    from oracle.unet import UNetOracle
    oracle = UNetOracle(n_classes, 1, 4, cf, seed=0)
    rng = np.random.RandomState(0)
    x = rng.randn(batch, dim, dim, 1).astype(np.float32)
    y = rng.randint(0, n_classes, size=(batch, dim, dim))
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
    return step


def cpu_train_baseline(args, steps, warmup=1):
    threads, usable, logical, physical = use_all_host_threads()
    step = cpu_reference_step_factory(args.cf, args.dim, args.classes)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    val = REF_BATCH * steps / dt
    sample = ("%d train steps (fwd + sparse-CE + bwd + Adam) of oracle/unet.py (torch-CPU fp32 restatement of the Keras "
              "step) on %d slices each after %d warm-up step(s), %.1f s; the GPU arm steps 32 slices - CPU slices/s does "
              "not depend on the batch (BASELINE.md 4.3: reduced batch, stated)" % (steps, REF_BATCH, warmup, dt))
    return {"value": val, "unit": "slices/s", "cores": threads, "kind": "port", "sample": sample,
            "host": {"usable_cpus": usable, "logical_cpus": logical, "physical_cores": physical}}, dt


def cpu_predict_baseline(args):
    """Reference-shaped CPU predict of one volume (oracle/predict_pipeline.py), bounded sample scaled to the volume."""
    import torch
    from oracle import predict_pipeline as pp
    from oracle.unet import UNetOracle
    from multiplanarunet_b200.interpolation import sample_random_views_with_angle_restriction
    from multiplanarunet_b200.sequences.isotrophic_live_view_sequence_2d import robust_scaler_stats
    threads, usable, logical, physical = use_all_host_threads()
    vol, _ = synthetic_volume(args.dim, args.classes, 1000, torch.device("cpu"))
    vol = vol.numpy()
    np.random.seed(0)
    views = sample_random_views_with_angle_restriction(6, 60)
    cen, scl = robust_scaler_stats(vol)
    bg = [float(np.percentile(vol[..., 0], 1))]
    oracle = UNetOracle(args.classes, 1, 4, args.cf, seed=0)
    W = np.random.RandomState(0).uniform(0.5, 1.5, (6, args.classes)).astype(np.float32)
    b = np.zeros(args.classes, np.float32)
    t0 = time.perf_counter()
    stages, total, sample = pp.time_predict_sample(vol, np.eye(4), views, args.dim, float(args.dim), bg, cen, scl,
                                                   lambda x: oracle.predict(x), W, b, n_classes=args.classes)
    spent = time.perf_counter() - t0
    return {"seconds_per_volume": total, "volumes_per_sec": 1.0 / total, "stages_seconds_per_volume": stages,
            "cores": threads, "kind": "port", "sample": sample + "; %.1f s of CPU work" % spent,
            "host": {"usable_cpus": usable, "logical_cpus": logical, "physical_cores": physical}}


def cpu_fusion_baseline(args, points=2 ** 21, epochs=1):
    """The fusion layer's train step restated on the CPU (oracle/fusion.py: float64 numpy GDL gradients + Adam),
    batch 2^17 as the reference, on a bounded number of points."""
    from oracle import fusion
    threads, usable, logical, physical = use_all_host_threads()
    rng = np.random.RandomState(0)
    V, C = 6, args.classes
    X = rng.rand(points, V, C).astype(np.float32)
    y = rng.randint(0, C, size=points).astype(np.uint8)
    W, b = np.ones((V, C)), np.zeros(C)
    th = np.concatenate([W.ravel(), b])
    m, v = np.zeros_like(th), np.zeros_like(th)
    t0 = time.perf_counter()
    t = 0
    for _ in range(epochs):
        for s in range(0, points, 2 ** 17):
            _, dW, db = fusion.gdl_loss_and_grads(X[s:s + 2 ** 17], y[s:s + 2 ** 17], th[:V * C].reshape(V, C), th[V * C:])
            t += 1
            th, m, v = fusion.adam_step(th, np.concatenate([dW.ravel(), db]), m, v, t, 1e-3)
    dt = time.perf_counter() - t0
    gbs = points * epochs * FUSION_BYTES_PER_POINT(V, C) / dt / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": "%d epoch(s) over %d points in batches of 2^17 (numpy float64 GDL gradients + Adam, "
                      "oracle/fusion.py), %.1f s" % (epochs, points, dt),
            "host": {"usable_cpus": usable, "logical_cpus": logical, "physical_cores": physical}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "train":
        base, dt = cpu_train_baseline(args, args.steps, min(args.warmup, 1))
        val, unit, metric = base["value"], "slices/s", "slices_per_sec"
        ms = 1e3 * dt / args.steps
        cfg = workload_config(args, REF_BATCH, "cpu")
    elif args.workload == "predict":
        base = cpu_predict_baseline(args)
        val, unit, metric = base["volumes_per_sec"], "volumes/s", "predict_volumes_per_sec"
        base["value"], base["unit"] = val, unit
        ms = 1e3 * base["seconds_per_volume"]
        cfg = predict_config(args, "cpu")
    else:
        base = cpu_fusion_baseline(args)
        val, unit, metric = base["value"], "GB/s", "fusion_train_gbytes_per_sec"
        ms = None
        cfg = fusion_config(args, "cpu", 1)
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "cpu_baseline": base,
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, batch, where):
    return {"workload": "6-view 2D U-Net train step on %dx%d oblique slices of a synthetic %d^3 volume" % (
        args.dim, args.dim, args.dim), "slices_per_step_per_gpu": batch, "complexity_factor": args.cf,
        "n_classes": args.classes, "n_views": 6, "optimizer": "Adam", "where": where,
        "l2_policy": "per-step working set (>10 GB of activations) far exceeds the 126 MB L2; no explicit flush"}


def predict_config(args, where):
    return {"workload": "mp predict of one synthetic %d^3 x 1ch volume: 6 views x %d planes of %dx%d, U-Net inference "
                        "(complexity_factor %g, %d classes) + nearest mapping + fusion + argmax" % (
                            args.dim, args.dim + 20, args.dim, args.dim, args.cf, args.classes), "where": where,
            "planes_per_inference_call": getattr(args, "predict_batch", None),
            "l2_policy": "every view streams >2 GB of activations; no explicit flush"}


def fusion_config(args, where, world):
    return {"workload": "train_fusion epochs over %d^3 x 6 views x %d classes mapped softmax points per GPU, batch 2^17 "
                        "per rank, shuffled every epoch" % (args.dim, args.classes), "where": where,
            "points_per_gpu": args.dim ** 3, "bytes_per_point": FUSION_BYTES_PER_POINT(6, args.classes),
            "all_reduce": ("36 doubles + point count per batch, exchanged inside the train-step kernel over NVLink "
                           "peer memory (NCCL only for the rendezvous)") if world > 1 else "none (single process)",
            "l2_policy": "2 GB of points per epoch >> 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------ GPU arms
def dist_setup():
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
    return world, rank, local, dev


def timed(fn, nwarm, nsteps, first, world, dev):
    """nwarm untimed + nsteps timed calls of fn(i), barrier + synchronize on both sides, max over ranks (ms)."""
    import torch
    import torch.distributed as dist
    from multiplanarunet_b200._C import lib
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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def bench_predict(args, model, views, dev, reps=2):
    """Config 3 through the public call chain: host numpy volume -> host uint8 label map."""
    import torch
    from multiplanarunet_b200.interpolation import ViewInterpolator
    from multiplanarunet_b200.sequences import IsotrophicLiveViewSequence2D, SyntheticImage
    from multiplanarunet_b200.utils.fusion.fuse_and_predict import predict_multi_view
    dim = args.dim
    vol, _ = synthetic_volume(dim, args.classes, 1000, dev)
    vol_host = vol.cpu().pin_memory()
    del vol
    # what loading an image does before the predict loop in the reference too (image_pair.py: bg value, scaler fit)
    image = SyntheticImage(vol_host.numpy(), None, np.eye(4), device=dev)
    seq = IsotrophicLiveViewSequence2D([image], views=views, sample_dim=dim, real_space_span=float(dim),
                                       n_classes=args.classes, is_validation=True)
    Wf = np.random.RandomState(0).uniform(0.5, 1.5, (6, args.classes)).astype(np.float32)
    bf = np.zeros(args.classes, np.float32)
    out_host = torch.empty(dim, dim, dim, dtype=torch.uint8).pin_memory()

    def once():
        t0 = time.perf_counter()
        # H2D of the volume: a fresh device-resident interpolator for this image
        image.interpolator = ViewInterpolator(vol_host, None, image.affine, bg_value=image.bg_value, device=dev)
        labels, _, _ = predict_multi_view(model, seq, image, views, Wf, bf, batch_size=model.max_batch)
        out_host.copy_(labels, non_blocking=False)   # D2H of the label map
        torch.cuda.synchronize()
        return time.perf_counter() - t0
    once()  # warm-up
    ts = [once() for _ in range(reps)]
    sec = float(np.median(ts))
    n_slices = 6 * (dim + 20)
    flops = n_slices * FWD_GFLOP_PER_SLICE.get(args.cf, 0.0) * 1e9 * (dim / 256.0) ** 2
    return {"volumes_per_sec": 1.0 / sec, "seconds_per_volume": sec, "h2d_bytes": int(vol_host.numel() * 4),
            "d2h_bytes": int(out_host.numel()), "slices": n_slices, "unet_algorithmic_tflops": flops / sec / 1e12,
            "foreground_fraction": float((out_host > 0).float().mean().item()),
            "config": predict_config(args, "B200")["workload"]}


def bench_train_fusion(args, world, rank, dev, epochs, warm_epochs=1):
    """Config 5: FusionModel.fit epochs over device-resident mapped-softmax points (one 256^3 volume per GPU)."""
    import torch
    from multiplanarunet_b200 import _C
    from multiplanarunet_b200._C import lib, check
    from multiplanarunet_b200.models import FusionModel
    V, C, N = 6, args.classes, args.dim ** 3
    g = torch.Generator(device=dev).manual_seed(2000 + rank)
    X = torch.rand(N, V, C, device=dev, generator=g)
    X = X / X.sum(-1, keepdim=True)
    y = torch.randint(0, C, (N,), device=dev, dtype=torch.uint8, generator=g)
    fm = FusionModel(V, C, device=dev)
    bs = 2 ** 17
    peer = fm.enable_peer_exchange() if world > 1 else False   # fused compute + NVLink peer-memory exchange kernel
    nb_all = (N + bs - 1) // bs

    def epoch(_):
        fm.fit(X, y, batch_size=bs, epochs=1, verbose=0, steps_per_epoch=nb_all if world > 1 else None)
    ms, launches = timed(epoch, warm_epochs, epochs, 0, world, dev)
    bytes_epoch = N * FUSION_BYTES_PER_POINT(V, C)
    gbs = world * bytes_epoch * epochs / (ms * 1e-3) / 1e9
    # the kernel alone: one un-shuffled pass, CUDA events on the launching stream
    acc = torch.zeros(V * C + C + 1, dtype=torch.float64, device=dev)
    st = _C.current_stream()

    def kern():
        check(lib.mpu_fusion_grad(_C.ptr(X), _C.ptr(y), ctypes.c_longlong(N), V, C, _C.ptr(fm.W), _C.ptr(fm.b),
                                  _C.ptr(acc), st), "mpu_fusion_grad")
    for _ in range(3):
        kern()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        kern()
    b.record()
    torch.cuda.synchronize()
    k_ms = a.elapsed_time(b) / 5
    hbm = load_peaks().get("hbm_gbs") or 6650.0
    return {"gbytes_per_sec": gbs, "points_per_sec": world * N * epochs / (ms * 1e-3), "ms_per_epoch": ms / epochs,
            "epochs": epochs, "batches_per_epoch": (N + bs - 1) // bs, "launches": int(launches),
            "kernel_ms_per_pass": k_ms, "kernel_gbytes_per_sec": bytes_epoch / k_ms / 1e6,
            "kernel_frac_of_measured_hbm": bytes_epoch / k_ms / 1e6 / hbm, "hbm_peak_gbs": hbm,
            "last_W_mean": float(fm.W.mean().item()),
            "exchange": ("fused into the train-step kernel over NVLink peer memory (no NCCL per batch)" if peer
                         else ("NCCL all-reduce per batch" if world > 1 else "none"))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "predict", "train_fusion"])
    ap.add_argument("--cf", type=float, default=2.0)
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--predict-batch", type=int, default=92,
                    help="planes per U-Net inference call of --workload predict (276 planes per view = 3 x 92: fuller "
                         "waves of the persistent GEMM kernels on the coarse levels than 32-plane calls)")
    ap.add_argument("--classes", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world, rank, local, dev = dist_setup()
    from multiplanarunet_b200 import _C
    from multiplanarunet_b200._C import lib, check
    from multiplanarunet_b200.interpolation import plane_basis_batch, sample_random_views_with_angle_restriction
    from multiplanarunet_b200.models import UNet
    from multiplanarunet_b200.sequences import IsotrophicLiveViewSequence2D, SyntheticImage
    lib.mpu_launch_count.restype = ctypes.c_longlong
    peaks = load_peaks()
    B, dim, K, W = args.batch, args.dim, args.steps, args.warmup
    np.random.seed(0)
    views = sample_random_views_with_angle_restriction(6, 60)

    # ---------------------------------------------------------------- alternative primary workloads
    if args.workload == "train_fusion":
        clocks = ClockSampler(local) if rank == 0 else None
        if clocks:
            clocks.start()
            time.sleep(0.3)
        r = bench_train_fusion(args, world, rank, dev, epochs=max(K, 3), warm_epochs=max(W // 2, 1))
        clock_info = clocks.stop() if clocks else None
        if rank == 0:
            base = None if (args.no_cpu_baseline or world > 1) else cpu_fusion_baseline(args)
            hbm = r["hbm_peak_gbs"]
            print(json.dumps({
                "metric": "fusion_train_gbytes_per_sec", "value": r["gbytes_per_sec"], "unit": "GB/s", "n_gpus": world,
                "steps": r["epochs"], "warmup": max(W // 2, 1), "ms_per_step": r["ms_per_epoch"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": fusion_config(args, "B200", world), "clocks": clock_info,
                "e2e": {"value": r["gbytes_per_sec"], "unit": "GB/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 8, "note": "points are produced on the device by predict_and_map; an epoch "
                        "through FusionModel.fit ends with one 8-byte read of the epoch loss"},
                "gpu_launches": r["launches"],
                "roofline": {"bound": "hbm", "achieved": r["kernel_gbytes_per_sec"], "peak": hbm, "unit": "GB/s",
                             "frac": r["kernel_frac_of_measured_hbm"], "traffic": None,
                             "kernels": "fusion_grad_kernel_t<6,%d> (one pass over %d points, %.3f ms)" % (
                                 args.classes, dim ** 3, r["kernel_ms_per_pass"]),
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback"},
                "cpu_baseline": base, "extras": r}))
        if world > 1:
            dist.destroy_process_group()
        return

    model = UNet(n_classes=args.classes, dim=dim, n_channels=1, complexity_factor=args.cf,
                 max_batch=args.predict_batch if args.workload == "predict" else B,
                 training=args.workload == "train", seed=0, device=dev)
    if args.workload == "predict":
        clocks = ClockSampler(local) if rank == 0 else None
        if clocks:
            clocks.start()
            time.sleep(0.3)
        r = bench_predict(args, model, views, dev, reps=max(min(K, 5), 2))
        t = torch.tensor([r["seconds_per_volume"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        clock_info = clocks.stop() if clocks else None
        if rank == 0:
            sec = float(t.item())
            base = None if (args.no_cpu_baseline or world > 1) else cpu_predict_baseline(args)
            peak = peaks.get("bf16_tflops_sustained") or 1400.0
            print(json.dumps({
                "metric": "predict_volumes_per_sec", "value": world / sec, "unit": "volumes/s", "n_gpus": world,
                "steps": max(min(K, 5), 2), "warmup": 1, "ms_per_step": 1e3 * sec, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": predict_config(args, "B200"), "clocks": clock_info,
                "e2e": {"value": world / sec, "unit": "volumes/s", "h2d_bytes_per_step": r["h2d_bytes"],
                        "d2h_bytes_per_step": r["d2h_bytes"]},
                "gpu_launches": None,
                "roofline": {"bound": "tensor", "achieved": r["unet_algorithmic_tflops"], "peak": peak,
                             "unit": "TFLOP/s", "frac": r["unet_algorithmic_tflops"] / peak, "traffic": None,
                             "kernels": "whole predict call (U-Net inference dominates: %d slices)" % r["slices"]},
                "cpu_baseline": base,
                "predict_vs_cpu": (base["seconds_per_volume"] / sec) if base else None, "extras": r}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- default: train step
    vol, labels = synthetic_volume(dim, args.classes, 1000 + rank, dev)
    image = SyntheticImage(vol.cpu().numpy(), labels.cpu().numpy(), np.eye(4), device=dev)
    del vol, labels
    model.optimizer.lr, model.optimizer.epsilon = 5e-5, 1e-8
    span = float(dim)
    # planes for every step are drawn on the host before the clock starts (view, offset ~ U(-span//2,
    # span//2), normal noise sd 0.1 as in sequences/isotrophic_live_view_sequence_2d.py:119-141)
    rng = np.random.RandomState(100 + rank)
    n_steps_total = W + K + 2
    planes = np.empty((n_steps_total, B, 10))
    for s in range(n_steps_total):
        planes[s, :, :9] = plane_basis_batch(views[rng.randint(0, 6, B)], rng.normal(scale=0.1, size=(B, 3))).reshape(B, 9)
        planes[s, :, 9] = rng.uniform(-(span // 2), span // 2, B)
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
        # and Adam applied range by range as each reduction completes
        model.forward_backward_overlapped(None, y_dev, None, input_packed=True, batch=B, fuse_adam=True)

    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
        time.sleep(0.3)
    ms_value, launches = timed(step_value, W, K, 0, world, dev)
    clock_info = clocks.stop() if clocks else None

    # ---- e2e: host (pinned) batches through the Keras-like public call model.train_on_batch(x, y, w)
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
        losses.append(model.train_on_batch(x_host, y_host, w_host))  # H2D of x / y / w inside, D2H of the loss

    ms_e2e, _ = timed(step_e2e, 2, K, 0, world, dev)
    h2d = x_host.numel() * 4 + y_host.numel() + w_host.numel() * 4
    d2h = 8

    # ---- roofline of the dominant kernels (tensor-core GEMMs), timed live with CUDA events
    # (every rank runs the extra step - it contains the gradient all-reduce - only rank 0 reads the timer)
    roofline = None
    R_PROF = 3  # serial steps averaged (a single step's sum moved by +-5 % between runs with the power-capped clock)
    check(lib.mpu_profile_gemm(1))
    for r_ in range(R_PROF):
        step_value(W + K + 1)
    gemm_ms, n_l = ctypes.c_double(), ctypes.c_int()
    check(lib.mpu_profile_gemm_read(ctypes.byref(gemm_ms), ctypes.byref(n_l)))
    check(lib.mpu_profile_gemm(0))
    gemm_ms.value /= R_PROF
    n_l.value //= R_PROF
    if rank == 0:
        # algorithmic FLOPs executed by the tensor-core GEMM kernels: 3x forward minus the first conv
        # (CUDA-core kernels, no input gradient) and the 1x1 head (fused softmax/CE kernel)
        non_gemm = {2.0: 0.106 + 0.059, 1.0: 0.075 + 0.042}.get(args.cf, 0.0)
        flops = 3.0 * (FWD_GFLOP_PER_SLICE.get(args.cf, 0.0) - non_gemm) * 1e9 * (dim / 256.0) ** 2 * B
        traffic, traffic_src = None, None
        for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", name)))
                if args.cf == 2.0 and dim == 256 and B == 32:
                    traffic, traffic_src = tr["gemm_dram_bytes_per_launch"], name
                break
            except Exception:
                continue
        peak = peaks.get("bf16_tflops_sustained")
        src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long step)"
        if not peak:
            peak, src = 1400.0, "fallback (B200_PROFILING.md sustained figure)"
        ach = flops / (gemm_ms.value * 1e-3) / 1e12 if gemm_ms.value > 0 else None
        roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": (ach / peak) if ach else None, "traffic": traffic,
                    "traffic_note": ("mean dram__bytes_read+write per GEMM launch, ncu pass over one train step "
                                     "(profiles/%s)" % traffic_src) if traffic else None,
                    "kernels": "mtgemm_fwd_kernel + mtgemm_wgrad_kernel (%d launches/step, %.2f ms/step)" % (
                        n_l.value, gemm_ms.value),
                    "algorithmic_flops_per_step": flops, "peak_source": src}

    # ---- the other BASELINE configurations and the throughput through UNet.fit (N=1 only, bounded)
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        extras = {}
        # (a) UNet.fit fed by the device rejection sampler + Elastic2D(p=1/3), prefetched on a side stream
        from multiplanarunet_b200.augmentation import Elastic2D
        seq_fit = IsotrophicLiveViewSequence2D([image], views=views, sample_dim=dim, real_space_span=span,
                                               n_classes=args.classes, batch_size=B, noise_sd=0.1,
                                               list_of_augmenters=[Elastic2D([0, 450], [20, 30], 0.333)])
        batches = seq_fit.prefetched(depth=2, rng=np.random.RandomState(5))
        try:
            model.fit(batches, steps_per_epoch=3, epochs=1, verbose=0)   # warm-up
            torch.cuda.synchronize()
            n_fit = max(K, 10)
            t0 = time.perf_counter()
            hist = model.fit(batches, steps_per_epoch=n_fit, epochs=1, verbose=0)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            extras["fit"] = {"slices_per_sec": B * n_fit / dt, "ms_per_step": 1e3 * dt / n_fit, "steps": n_fit,
                             "loss": hist["loss"][-1],
                             "what": "UNet.fit over IsotrophicLiveViewSequence2D.prefetched(): per batch 320 candidate "
                                     "planes probed on the device, reference accept rules on the host, accepted planes "
                                     "sampled + scaled, Elastic2D(alpha [0,450], sigma [20,30], apply_prob 0.333)"}
        finally:
            batches.close()
        # (b) config 3
        model_inf = model  # same weights; inference uses the moving statistics
        extras["predict"] = bench_predict(args, model_inf, views, dev, reps=2)
        # (c) config 5 at N = 1
        del image, seq_fit
        torch.cuda.empty_cache()
        extras["train_fusion"] = bench_train_fusion(args, 1, 0, dev, epochs=5)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = cpu_train_baseline(args, steps=3, warmup=1)
        if extras is not None:
            pc = cpu_predict_baseline(args)
            extras["predict"]["cpu_baseline"] = pc
            extras["predict"]["predict_vs_cpu"] = pc["seconds_per_volume"] / extras["predict"]["seconds_per_volume"]
            fc = cpu_fusion_baseline(args)
            extras["train_fusion"]["cpu_baseline"] = fc
            extras["train_fusion"]["vs_cpu"] = extras["train_fusion"]["gbytes_per_sec"] / fc["value"]

    if rank == 0:
        out = {
            "metric": "slices_per_sec", "value": world * B * K / (ms_value * 1e-3), "unit": "slices/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_value / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, B, "B200"),
            "clocks": clock_info,
            "e2e": {"value": world * B * K / (ms_e2e * 1e-3), "unit": "slices/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "last_loss": losses[-1] if losses else None,
                    "call": "UNet.train_on_batch(x, y, sample_weight) with pinned host tensors"},
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
