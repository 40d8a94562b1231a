"""GPU bring-up harness for the tcgen05 multi-tap GEMM kernels (run on the B200 box via gpurun).

Each case runs in its own subprocess under a timeout so a trapped kernel cannot take the rest down.
  python tests/bringup_gemm.py            # driver: run all cases, write gpurun_out/bringup_gemm.log
  python tests/bringup_gemm.py --case X   # one case
The reference is torch fp32 conv2d on the same bf16-rounded operands.
"""
import argparse
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pad_nhwc(x, cphys):
    import torch
    B, H, W, C = x.shape
    out = torch.zeros(B, H + 2, W + 2, cphys, dtype=torch.bfloat16, device=x.device)
    out[:, 1:-1, 1:-1, :C] = x.to(torch.bfloat16)
    return out


def conv_case(B, H, W, Cin, Cout, BN, cin_phys=None, cout_phys=None, relu=True, bias=True, mask=False,
              two_src=False, seed=0):
    import torch
    import torch.nn.functional as F
    from multiplanarunet_b200._C import lib, check, ptr, int_array
    torch.manual_seed(seed)
    dev = "cuda"
    cin_phys = cin_phys or ((Cin + 7) // 8 * 8)
    cout_phys = cout_phys or ((Cout + 7) // 8 * 8)
    Hp, Wp = H + 2, W + 2
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    xp = pad_nhwc(x, cin_phys)
    srcs = [(x, xp, Cin, cin_phys)]
    if two_src:
        x2 = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
        srcs.append((x2, pad_nhwc(x2, cin_phys), Cin, cin_phys))
    ktot = sum(s[3] for s in srcs)
    cin_tot = sum(s[2] for s in srcs)
    w = (torch.randn(3, 3, cin_tot, Cout, device=dev) / (3.0 * cin_tot ** 0.5)).to(torch.bfloat16)
    # kernel layout [tap][cout_phys][ktot]
    wk = torch.zeros(9, cout_phys, ktot, dtype=torch.bfloat16, device=dev)
    k0 = 0
    c0 = 0
    for (_, _, c, cp) in srcs:
        wk[:, :Cout, k0:k0 + c] = w[:, :, c0:c0 + c, :].permute(0, 1, 3, 2).reshape(9, Cout, c)
        k0 += cp
        c0 += c
    b = torch.randn(cout_phys, device=dev) * 0.1
    b[Cout:] = 0
    mk = None
    if mask:
        mk = pad_nhwc(torch.randn(B, H, W, Cout, device=dev), cout_phys)
    out = torch.zeros(B, Hp, Wp, cout_phys, dtype=torch.bfloat16, device=dev)
    taps_off = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
    taps_w = list(range(9))
    A1 = srcs[1][1] if two_src else None
    rc = lib.mpu_mtgemm_fwd(
        ptr(xp), ctypes.c_longlong(B * Hp * Wp), cin_phys, cin_phys,
        ptr(A1), ctypes.c_longlong(B * Hp * Wp if two_src else 0), cin_phys if two_src else 0,
        cin_phys if two_src else 0,
        ptr(wk), 9, cout_phys, ktot, 9, int_array(taps_off), int_array(taps_w), B * Hp * Wp,
        Hp, Wp, Hp, Wp, 1, 0, 0, ptr(out), cout_phys, ptr(b if bias else None), ptr(mk), cout_phys,
        1 if relu else 0, ctypes.c_void_p(0))
    check(rc, "mpu_mtgemm_fwd")
    torch.cuda.synchronize()
    xin = torch.cat([s[0] for s in srcs], dim=-1).float().permute(0, 3, 1, 2)
    ref = F.conv2d(xin, w.float().permute(3, 2, 0, 1), padding=1)
    if bias:
        ref = ref + b[:Cout].view(1, -1, 1, 1)
    if relu:
        ref = ref.clamp_min(0)
    ref = ref.permute(0, 2, 3, 1)
    if mask:
        ref = ref * (mk[:, 1:-1, 1:-1, :Cout].float() > 0)
    got = out[:, 1:-1, 1:-1, :Cout].float()
    err = (got - ref).abs()
    tol = 0.02 + 0.01 * ref.abs()
    bad = (err > tol).float().mean().item()
    bm = torch.ones_like(out, dtype=torch.bool)
    bm[:, 1:-1, 1:-1, :] = False
    border = out[bm].float().abs().sum()
    padc = out[..., Cout:].float().abs().sum().item()
    print("  max_err=%.4g mean_err=%.4g frac_bad=%.4g ref_absmean=%.4g border_sum=%.4g padch_sum=%.4g" %
          (err.max().item(), err.mean().item(), bad, ref.abs().mean().item(), border.item(), padc))
    if bad > 0:
        idx = (err > tol).nonzero()[:5]
        for i in idx:
            i = tuple(i.tolist())
            print("   bad at", i, "got", got[i].item(), "ref", ref[i].item())
        # error structure by channel and by x position
        print("   bad by channel(first 16):", (err > tol).float().mean(dim=(0, 1, 2))[:16].tolist())
        print("   bad by x(first 16):", (err > tol).float().mean(dim=(0, 1, 3))[:16].tolist())
    return bad == 0 and border.item() == 0


def perf_case(B, H, W, Cin, Cout, iters=10, two_src=False):
    """Device-timed 3x3 conv at U-Net-like sizes; prints TFLOP/s on algorithmic (unpadded) FLOPs."""
    import torch
    from multiplanarunet_b200._C import lib, check, ptr, int_array
    dev = "cuda"
    cin_phys = (Cin + 7) // 8 * 8
    cout_phys = (Cout + 7) // 8 * 8
    Hp, Wp = H + 2, W + 2
    nsrc = 2 if two_src else 1
    xs = [pad_nhwc(torch.randn(B, H, W, Cin, device=dev), cin_phys) for _ in range(nsrc)]
    wk = (torch.randn(9, cout_phys, nsrc * cin_phys, device=dev) * 0.05).to(torch.bfloat16)
    out = torch.zeros(B, Hp, Wp, cout_phys, dtype=torch.bfloat16, device=dev)
    taps_off = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]

    def run():
        rc = lib.mpu_mtgemm_fwd(
            ptr(xs[0]), ctypes.c_longlong(B * Hp * Wp), cin_phys, cin_phys,
            ptr(xs[1] if two_src else None), ctypes.c_longlong(B * Hp * Wp if two_src else 0),
            cin_phys if two_src else 0, cin_phys if two_src else 0,
            ptr(wk), 9, cout_phys, nsrc * cin_phys, 9, int_array(taps_off), int_array(list(range(9))),
            B * Hp * Wp, Hp, Wp, Hp, Wp, 1, 0, 0, ptr(out), cout_phys, ptr(None), ptr(None), 0, 1,
            ctypes.c_void_p(0))
        check(rc, "mpu_mtgemm_fwd")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    fl = 2.0 * B * H * W * 9 * nsrc * Cin * Cout
    print("  perf B=%d %dx%d %d->%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (B, H, W, nsrc * Cin, Cout, ms, fl / ms / 1e9))
    # role-stall profile of CTA 0
    cnt = torch.zeros(8, dtype=torch.int64, device=dev)
    lib.mpu_debug_set_fwd_profile(ctypes.c_void_p(cnt.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.mpu_debug_set_fwd_profile(ctypes.c_void_p(0))
    c = cnt.cpu().tolist()
    tot = max(c[7], 1)
    print("   CTA0 cycles %d | producer wait a_empty %.0f%% b_empty %.0f%% | mma wait a_full %.0f%% b_full %.0f%% "
          "tmem_empty %.0f%% | epilogue wait tmem_full %.0f%% busy %.0f%%" %
          (tot, 100 * c[0] / tot, 100 * c[1] / tot, 100 * c[2] / tot, 100 * c[3] / tot, 100 * c[4] / tot,
           100 * c[5] / tot, 100 * c[6] / tot))
    return True


def perf_wgrad_case(B, H, W, Cin, Cout, iters=10):
    import torch
    from multiplanarunet_b200._C import lib, check, ptr, int_array
    dev = "cuda"
    cin_phys = (Cin + 7) // 8 * 8
    cout_phys = (Cout + 7) // 8 * 8
    Hp, Wp = H + 2, W + 2
    xp = pad_nhwc(torch.randn(B, H, W, Cin, device=dev), cin_phys)
    dyp = pad_nhwc(torch.randn(B, H, W, Cout, device=dev), cout_phys)
    dW = torch.zeros(9, cout_phys, cin_phys, dtype=torch.float32, device=dev)
    taps_off = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
    rows = B * Hp * Wp

    def run():
        rc = lib.mpu_mtgemm_wgrad(
            ptr(xp), ctypes.c_longlong(rows), cin_phys, cin_phys, ptr(dyp), ctypes.c_longlong(rows),
            cout_phys, cout_phys, 9, int_array(taps_off), None, int_array(list(range(9))),
            ctypes.c_longlong(rows), 0, ptr(dW), cin_phys, cout_phys, 0, ctypes.c_void_p(0))
        check(rc, "mpu_mtgemm_wgrad")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    fl = 2.0 * B * H * W * 9 * Cin * Cout
    print("  wgrad perf B=%d %dx%d %d->%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (B, H, W, Cin, Cout, ms, fl / ms / 1e9))
    # role timings of CTA 0 (profiling instantiation)
    cnt = torch.zeros(8, dtype=torch.int64, device=dev)
    lib.mpu_debug_set_fwd_profile(ctypes.c_void_p(cnt.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.mpu_debug_set_fwd_profile(ctypes.c_void_p(0))
    c = cnt.cpu().tolist()
    tot = max(c[4], 1)
    print("   CTA0 cycles %d (prologue %d, MMAs retired at %d, drain %d) | %d K blocks, %.0f cycles each | producer waits "
          "empty %.0f%% | mma waits full %.0f%%" %
          (tot, c[5], c[2], c[3], c[6], (c[2] - c[5]) / max(c[6], 1), 100 * c[0] / tot, 100 * c[1] / tot))
    return True


def upconv_case(B, h, w_, Cin, Cout, BN, seed=0):
    """nearest-2x upsample + 2x2 SAME conv via 4 phase-collapsed multi-tap GEMMs."""
    import torch
    import torch.nn.functional as F
    from multiplanarunet_b200._C import lib, check, ptr, int_array
    torch.manual_seed(seed)
    dev = "cuda"
    cin_phys = (Cin + 7) // 8 * 8
    cout_phys = (Cout + 7) // 8 * 8
    hp, wp = h + 2, w_ + 2
    H, W = 2 * h, 2 * w_
    Hp, Wp = H + 2, W + 2
    x = torch.randn(B, h, w_, Cin, device=dev).to(torch.bfloat16)
    xp = pad_nhwc(x, cin_phys)
    wt = (torch.randn(2, 2, Cin, Cout, device=dev) / (2.0 * Cin ** 0.5)).float()
    # collapsed weights per (phase, tap): list of (phase a,b, tap di,dj, weight)
    pairs = []
    for a in range(2):
        for b_ in range(2):
            acc = {}
            for dy in range(2):
                for dx in range(2):
                    di, dj = (a + dy) >> 1, (b_ + dx) >> 1
                    acc[(di, dj)] = acc.get((di, dj), 0) + wt[dy, dx]
            for (di, dj), wsum in sorted(acc.items()):
                pairs.append((a, b_, di, dj, wsum))
    wk = torch.zeros(len(pairs), cout_phys, cin_phys, dtype=torch.bfloat16, device=dev)
    for i, (_, _, _, _, ws) in enumerate(pairs):
        wk[i, :Cout, :Cin] = ws.t().to(torch.bfloat16)
    bias = torch.randn(cout_phys, device=dev) * 0.1
    bias[Cout:] = 0
    out = torch.zeros(B, Hp, Wp, cout_phys, dtype=torch.bfloat16, device=dev)
    for a in range(2):
        for b_ in range(2):
            idx = [i for i, p in enumerate(pairs) if p[0] == a and p[1] == b_]
            offs = [pairs[i][2] * wp + pairs[i][3] for i in idx]
            rc = lib.mpu_mtgemm_fwd(
                ptr(xp), ctypes.c_longlong(B * hp * wp), cin_phys, cin_phys, ptr(None),
                ctypes.c_longlong(0), 0, 0, ptr(wk), len(pairs), cout_phys, cin_phys, len(idx),
                int_array(offs), int_array(idx), B * hp * wp, hp, wp, Hp, Wp, 2, a, b_, ptr(out),
                cout_phys, ptr(bias), ptr(None), 0, 1, ctypes.c_void_p(0))
            check(rc, "mpu_mtgemm_fwd(upconv)")
    torch.cuda.synchronize()
    # reference with the same collapsed bf16 weights (so only accumulation order differs)
    xin = x.float()
    xpad = F.pad(xin, (0, 0, 0, 1, 0, 1))  # pad w and h at the end by 1
    ref = torch.zeros(B, H, W, Cout, device=dev)
    for i, (a, b_, di, dj, _) in enumerate(pairs):
        contrib = xpad[:, di:di + h, dj:dj + w_, :] @ wk[i, :Cout, :Cin].float().t()
        ref[:, a::2, b_::2, :] += contrib
    ref = (ref + bias[:Cout]).clamp_min(0)
    # and the plain definition (fp32 weights) for information
    up = xin.permute(0, 3, 1, 2).repeat_interleave(2, 2).repeat_interleave(2, 3)
    up = F.pad(up, (0, 1, 0, 1))
    ref2 = F.conv2d(up, wt.permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
    ref2 = (ref2 + bias[:Cout]).clamp_min(0)
    got = out[:, 1:-1, 1:-1, :Cout].float()
    err = (got - ref).abs()
    tol = 0.02 + 0.01 * ref.abs()
    bad = (err > tol).float().mean().item()
    print("  max_err=%.4g frac_bad=%.4g | vs plain-definition max_err=%.4g" %
          (err.max().item(), bad, (got - ref2).abs().max().item()))
    return bad == 0


def wgrad_case(B, H, W, Cin, Cout, BN, G=3, splits=4, variant=None, seed=0):
    import torch
    import torch.nn.functional as F
    from multiplanarunet_b200._C import lib, check, ptr, int_array
    torch.manual_seed(seed)
    dev = "cuda"
    cin_phys = (Cin + 7) // 8 * 8
    cout_phys = (Cout + 7) // 8 * 8
    Hp, Wp = H + 2, W + 2
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16)
    xp = pad_nhwc(x, cin_phys)
    dyp = pad_nhwc(dy, cout_phys)
    dW = torch.zeros(9, cout_phys, cin_phys, dtype=torch.float32, device=dev)
    taps_off = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
    v = variant or (0, 0, 0, 0, 0)
    rows = B * Hp * Wp
    rc = lib.mpu_mtgemm_wgrad(
        ptr(xp), ctypes.c_longlong(rows), cin_phys, cin_phys, ptr(dyp), ctypes.c_longlong(rows),
        cout_phys, cout_phys, 9, int_array(taps_off), int_array([0] * 9), int_array(list(range(9))),
        ctypes.c_longlong(rows), splits, ptr(dW), cin_phys, cout_phys, 0, ctypes.c_void_p(0))
    check(rc, "mpu_mtgemm_wgrad")
    torch.cuda.synchronize()
    xin = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wref = torch.zeros(Cout, Cin, 3, 3, device=dev, requires_grad=True)
    y = F.conv2d(xin, wref, padding=1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = wref.grad.permute(2, 3, 0, 1).reshape(9, Cout, Cin)  # [tap][co][ci]
    got = dW[:, :Cout, :Cin]
    err = (got - ref).abs()
    scale = ref.abs().mean().item()
    bad = (err > 0.02 * scale + 0.01 * ref.abs()).float().mean().item()
    padsum = dW.abs().sum().item() - got.abs().sum().item()
    print("  variant=%s max_err=%.4g ref_absmean=%.4g frac_bad=%.4g pad_sum=%.4g" %
          (str(v), err.max().item(), scale, bad, padsum))
    if bad > 0:
        print("   bad by tap:", (err > 0.02 * scale + 0.01 * ref.abs()).float().mean(dim=(1, 2)).tolist())
        print("   bad by ci(first 16):",
              (err > 0.02 * scale + 0.01 * ref.abs()).float().mean(dim=(0, 1))[:16].tolist())
        print("   bad by co(first 16):",
              (err > 0.02 * scale + 0.01 * ref.abs()).float().mean(dim=(0, 2))[:16].tolist())
        print("   sample got/ref:", got[4, 0, :4].tolist(), ref[4, 0, :4].tolist())
    return bad == 0


CASES = {
    "perf_L0": (perf_case, dict(B=32, H=256, W=256, Cin=90, Cout=90)),
    "perf_L0cat": (perf_case, dict(B=32, H=256, W=256, Cin=90, Cout=90, two_src=True)),
    "perf_L1": (perf_case, dict(B=32, H=128, W=128, Cin=181, Cout=181)),
    "perf_L2": (perf_case, dict(B=32, H=64, W=64, Cin=362, Cout=362)),
    "perf_L3": (perf_case, dict(B=32, H=32, W=32, Cin=724, Cout=724)),
    "wperf_L0": (perf_wgrad_case, dict(B=32, H=256, W=256, Cin=90, Cout=90)),
    "wperf_L1": (perf_wgrad_case, dict(B=32, H=128, W=128, Cin=181, Cout=181)),
    "wperf_L2": (perf_wgrad_case, dict(B=32, H=64, W=64, Cin=362, Cout=362)),
    "wperf_L3": (perf_wgrad_case, dict(B=32, H=32, W=32, Cin=724, Cout=724)),
    "wperf_L4": (perf_wgrad_case, dict(B=32, H=16, W=16, Cin=1448, Cout=1448)),
    "perf_L4": (perf_case, dict(B=32, H=16, W=16, Cin=1448, Cout=1448)),
    # name: (fn, kwargs)
    "fwd_small_64": (conv_case, dict(B=1, H=16, W=16, Cin=64, Cout=64, BN=64, relu=False, bias=False)),
    "fwd_small_64_relu_bias": (conv_case, dict(B=2, H=16, W=16, Cin=64, Cout=64, BN=64)),
    "fwd_c96_n96": (conv_case, dict(B=2, H=32, W=32, Cin=90, Cout=90, BN=96, cin_phys=96, cout_phys=96)),
    "fwd_c192_n192_mask": (conv_case, dict(B=2, H=32, W=32, Cin=181, Cout=181, BN=192, cin_phys=192,
                                           cout_phys=192, mask=True, relu=False, bias=False)),
    "fwd_two_src": (conv_case, dict(B=2, H=16, W=16, Cin=90, Cout=90, BN=96, cin_phys=96, cout_phys=96,
                                    two_src=True)),
    "fwd_big_n256": (conv_case, dict(B=4, H=64, W=64, Cin=362, Cout=362, BN=256, cin_phys=384,
                                     cout_phys=384)),
    "fwd_cin8": (conv_case, dict(B=2, H=32, W=32, Cin=1, Cout=90, BN=96, cin_phys=8, cout_phys=96)),
    "upconv": (upconv_case, dict(B=2, h=8, w_=8, Cin=128, Cout=64, BN=64)),
    "upconv_odd": (upconv_case, dict(B=2, h=16, w_=16, Cin=181, Cout=90, BN=96)),
    "wgrad_small": (wgrad_case, dict(B=1, H=16, W=16, Cin=128, Cout=64, BN=64, G=3, splits=1)),
    "wgrad_split": (wgrad_case, dict(B=2, H=32, W=32, Cin=128, Cout=128, BN=128, G=3, splits=8)),
    "wgrad_c96": (wgrad_case, dict(B=2, H=32, W=32, Cin=90, Cout=90, BN=96, G=3, splits=4)),
    "wgrad_n256": (wgrad_case, dict(B=2, H=16, W=16, Cin=362, Cout=362, BN=256, G=2, splits=2)),
    # channel counts of the deep levels at complexity_factor 2 (the benchmark configuration)
    "fwd_724": (conv_case, dict(B=2, H=32, W=32, Cin=724, Cout=724, BN=0)),
    "fwd_1448_16x16": (conv_case, dict(B=2, H=16, W=16, Cin=1448, Cout=1448, BN=0)),
    "fwd_two_src_724": (conv_case, dict(B=2, H=32, W=32, Cin=724, Cout=724, BN=0, two_src=True)),
    "fwd_1448_to_724_mask": (conv_case, dict(B=2, H=32, W=32, Cin=1448, Cout=724, BN=0, mask=True, relu=False,
                                             bias=False)),
    "upconv_1448_to_724": (upconv_case, dict(B=2, h=16, w_=16, Cin=1448, Cout=724, BN=0)),
    "upconv_181_to_90": (upconv_case, dict(B=1, h=128, w_=128, Cin=181, Cout=90, BN=0)),
    "wgrad_724": (wgrad_case, dict(B=2, H=32, W=32, Cin=724, Cout=724, BN=0, G=3, splits=0)),
    "wgrad_1448_16x16": (wgrad_case, dict(B=2, H=16, W=16, Cin=1448, Cout=1448, BN=0, G=3, splits=0)),
    "wgrad_1448_to_724": (wgrad_case, dict(B=2, H=32, W=32, Cin=1448, Cout=724, BN=0, G=3, splits=0)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--only", default=None, help="comma-separated subset for the driver")
    args = ap.parse_args()
    if args.case:
        fn, kw = CASES[args.case]
        ok = fn(**kw)
        print("CASE %s: %s" % (args.case, "PASS" if ok else "FAIL"))
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "bringup_gemm.log"), "w")
    names = list(CASES) if not args.only else args.only.split(",")
    npass = 0
    for name in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name],
                               capture_output=True, text=True, timeout=180)
            out = r.stdout + r.stderr[-3000:]
            status = "PASS" if r.returncode == 0 else "FAIL(rc=%d)" % r.returncode
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or "") + "\nTIMEOUT"
            status = "TIMEOUT"
        npass += status == "PASS"
        msg = "=== %s: %s\n%s\n" % (name, status, out)
        print(msg)
        log.write(msg)
        log.flush()
    print("SUMMARY: %d/%d passed" % (npass, len(names)))
    log.write("SUMMARY: %d/%d passed\n" % (npass, len(names)))


if __name__ == "__main__":
    main()
