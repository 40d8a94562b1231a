"""GPU bring-up for the volume kernels (sampler, map+fuse, fusion training) vs the numpy oracle.
Run on the B200 box via gpurun:  python tests/bringup_volume.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sampler_cases():
    import torch
    from multiplanarunet_b200.interpolation import ViewInterpolator, plane_basis
    from oracle import sampler
    rng = np.random.RandomState(0)
    ok = True
    for name, shape, affine in [
        ("iso", (24, 20, 28), np.eye(4)),
        ("aniso", (24, 20, 28), np.diag([1.0, 0.5, 2.0, 1.0])),
        ("rotated", (20, 20, 20), None),
    ]:
        if affine is None:
            th = 0.3
            R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
            affine = np.eye(4)
            affine[:3, :3] = R.dot(np.diag([1.0, 1.5, 0.75]))
        C = 2
        vol = rng.randn(*shape, C).astype(np.float32)
        lab = rng.randint(0, 4, size=shape).astype(np.uint8)
        bg = [-1.5, 0.25]
        center, scale = [0.1, -0.2], [1.3, 0.7]
        vi = ViewInterpolator(vol, lab, affine, bg_value=bg, bg_class=0)
        pix = np.linalg.norm(affine[:3, :3], axis=0)
        dim, span = 32, 30
        nbad_im = nbad_lab = 0
        maxdiff = 0.0
        total = 0
        for view in [(0.3, 0.5, 0.8), (0, 0, 1), (1, 0, 0), (-0.7, 0.1, 0.2), (0.05, 0.1, 0.99)]:
            basis = plane_basis(view, 0.)
            offs = np.array([-13.2, 0.0, 5.5, 11.9])
            im, lb = vi.sample_planes(basis, offs, dim, span, center=center, scale=scale)
            im, lb = im.cpu().numpy(), lb.cpu().numpy()
            for k, off in enumerate(offs):
                ro, rl = sampler.sample_plane(vol, lab, pix, basis, dim, span, off, bg, 0, center, scale,
                                              rot_mat=vi.rot_mat)
                nbad_im += int((im[k] != ro).sum())
                nbad_lab += int((lb[k] != rl).sum())
                maxdiff = max(maxdiff, float(np.abs(im[k] - ro).max()))
                total += ro.size
        print("sampler[%s]: image mismatches %d/%d (max abs diff %.3g), label mismatches %d" %
              (name, nbad_im, total, maxdiff, nbad_lab))
        ok = ok and nbad_lab == 0 and maxdiff < 1e-5
        # padded bf16 output
        cpad = 8
        pad = torch.zeros(2 * (dim + 2) * (dim + 2), cpad, dtype=torch.bfloat16, device="cuda")
        basis = plane_basis((0.3, 0.5, 0.8), 0.)
        im, _ = vi.sample_planes(basis, [0.0, 3.0], dim, span, center=center, scale=scale, out_padded=pad,
                                 cpad=cpad)
        got = pad.view(2, dim + 2, dim + 2, cpad)[:, 1:-1, 1:-1, :C].float()
        exp = im.to(torch.bfloat16).float()
        border = pad.view(2, dim + 2, dim + 2, cpad).float().abs().sum() - got.abs().sum()
        print("   padded bf16 input: equal=%s border=%.3g" % (bool((got == exp).all()), border.item()))
        ok = ok and bool((got == exp).all())
    return ok


def mapfuse_case():
    import torch
    from multiplanarunet_b200.interpolation import plane_basis, view_offsets
    from multiplanarunet_b200.utils.fusion.fuse_and_predict import _map_fuse
    from oracle import fusion
    rng = np.random.RandomState(1)
    ok = True
    for name, shape, affine in [("iso", (24, 20, 28), np.eye(3)), ("aniso", (20, 24, 16), np.diag([1.0, 0.8, 1.7]))]:
        dim, span, n, C, V = 32, 30, 40, 5, 3
        views = [(0.3, 0.5, 0.8), (0, 0, 1), (-0.7, 0.1, 0.2)]
        preds, grids, ibs = [], [], []
        g = np.linspace(-(span // 2), span // 2, dim)
        for v in views:
            p = rng.rand(n, dim, dim, C).astype(np.float32)
            p /= p.sum(-1, keepdims=True)
            preds.append(p)
            grids.append((g, g, view_offsets(dim, span, n)))
            ibs.append(np.linalg.inv(plane_basis(v, 0.)))
        W = rng.uniform(0.5, 1.5, size=(V, C)).astype(np.float32)
        b = (0.1 * rng.randn(C)).astype(np.float32)
        labels, probs, combined = _map_fuse([torch.as_tensor(p).cuda() for p in preds], grids, ibs, shape,
                                            affine, W, b, want_probs=True, want_combined=True)
        vg = fusion.voxel_grid_real_space(shape, affine)
        comb_ref = np.stack([fusion.map_real_space_pred(np.moveaxis(p, 0, 2), gr, ib, vg)
                             for p, gr, ib in zip(preds, grids, ibs)])
        probs_ref, labels_ref = fusion.merge_views(comb_ref, W, b)
        cm = int((combined.cpu().numpy() != comb_ref).sum())
        lm = int((labels.cpu().numpy() != labels_ref).sum())
        pd = float(np.abs(probs.cpu().numpy() - probs_ref).max())
        print("map_fuse[%s]: mapped mismatches %d/%d, label mismatches %d/%d, probs max diff %.3g, oob frac %.3f" %
              (name, cm, comb_ref.size, lm, labels_ref.size, pd, float((comb_ref[..., 0] == 1).mean())))
        ok = ok and cm == 0 and lm == 0 and pd < 1e-6
        labels2, _, _ = _map_fuse([torch.as_tensor(p).cuda() for p in preds], grids, ibs, shape, affine,
                                  sum_fusion=True)
        _, lref2 = fusion.merge_views(comb_ref, sum_fusion=True)
        lm2 = int((labels2.cpu().numpy() != lref2).sum())
        print("   sum_fusion label mismatches %d" % lm2)
        ok = ok and lm2 == 0
    return ok


def fusion_train_case():
    import ctypes
    import torch
    from multiplanarunet_b200 import _C
    from multiplanarunet_b200.models import FusionModel
    from oracle import fusion
    rng = np.random.RandomState(2)
    N, V, C = 20000, 6, 5
    X = rng.rand(N, V, C).astype(np.float32)
    X /= X.sum(-1, keepdims=True)
    y = rng.randint(0, C, size=N).astype(np.uint8)
    fm = FusionModel(V, C)
    W0 = rng.uniform(0.5, 1.5, size=(V, C)).astype(np.float32)
    b0 = (0.1 * rng.randn(C)).astype(np.float32)
    fm.set_weights([W0, b0])
    Xd, yd = torch.as_tensor(X).cuda(), torch.as_tensor(y).cuda()
    ok = True
    lref, dW, db = fusion.gdl_loss_and_grads(X, y, W0, b0)
    # gradient sums of the kernel (contiguous rows and through a permutation index) against the float64 oracle
    perm = torch.randperm(N, device="cuda")
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(N, device="cuda")
    for tag, Xk, yk, idx in (("rows", Xd, yd, None), ("indexed", Xd[perm].contiguous(), yd[perm].contiguous(), inv)):
        acc = torch.zeros(V * C + C + 1, dtype=torch.float64, device="cuda")
        _C.check(_C.lib.mpu_fusion_grad_indexed(_C.ptr(Xk), _C.ptr(yk), _C.ptr(idx), ctypes.c_longlong(N), V, C,
                                                _C.ptr(fm.W), _C.ptr(fm.b), _C.ptr(acc), _C.current_stream()))
        a = acc.cpu().numpy()
        gW = a[:V * C].reshape(V, C) / N + 1e-6 * 2 * W0 / W0.size
        gb = a[V * C:V * C + C] / N + 1e-6 * 2 * b0 / b0.size
        eW, eb = np.abs(gW - dW).max() / np.abs(dW).max(), np.abs(gb - db).max() / np.abs(db).max()
        el = abs(a[-1] / N + 1e-6 * (W0 ** 2).mean() + 1e-6 * (b0 ** 2).mean() - lref)
        print("fusion grads[%s]: dW rel err %.3g  db rel err %.3g  loss err %.3g" % (tag, eW, eb, el))
        ok = ok and eW < 1e-3 and eb < 1e-3 and el < 1e-5
    # one fused train step (gradient + Adam in one launch) against the oracle's Adam
    loss = fm.train_on_batch(Xd, yd)
    torch.cuda.synchronize()
    th, m, v = np.concatenate([W0.ravel(), b0]).astype(np.float64), np.zeros(V * C + C), np.zeros(V * C + C)
    th, m, v = fusion.adam_step(th, np.concatenate([dW.ravel(), db]), m, v, 1, 1e-3)
    W1, b1 = fm.get_weights()
    err = max(np.abs(W1.ravel() - th[:V * C]).max(), np.abs(b1.ravel() - th[V * C:]).max())
    print("fusion fused step: max |param - oracle| = %.3g (step size 1e-3), loss %.6f (oracle, no reg, %.6f)" % (
        err, float(loss), lref - 1e-6 * (W0 ** 2).mean() - 1e-6 * (b0 ** 2).mean()))
    ok = ok and err < 2e-5 and float(fm._accum.abs().max()) == 0.0 and int(fm._counter[:4].view(torch.int32).item()) == 0
    # a few shuffled epochs reduce the loss; evaluate() agrees with the last epoch's scale
    hist = fm.fit(Xd, yd, batch_size=4096, epochs=3)
    ev = fm.evaluate(Xd, yd)
    print("fusion fit losses:", ["%.5f" % h for h in hist], "evaluate %.5f" % ev)
    return ok and hist[-1] <= hist[0] and abs(ev - hist[-1]) < 0.05


if __name__ == "__main__":
    res = {}
    for name, fn in [("sampler", sampler_cases), ("mapfuse", mapfuse_case), ("fusion_train", fusion_train_case)]:
        try:
            res[name] = bool(fn())
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            res[name] = False
    print("RESULT:", res)
