"""Seeded inputs shared by oracle/make_golden.py (which ran the real reference on them) and the tests."""
import numpy as np

SAMPLER_CASES = [
    dict(name="iso", shape=(18, 14, 22), pix=(1.0, 1.0, 1.0), dim=24, span=22,
         views=[(0.3, 0.5, 0.8), (0, 0, 1), (1, 0, 0), (-0.7, 0.1, 0.2)], offsets=[-9.3, 0.0, 4.5, 10.9], seed=11),
    dict(name="aniso", shape=(16, 20, 12), pix=(1.0, 0.5, 2.0), dim=20, span=24,
         views=[(0.05, 0.1, 0.99), (0.5, -0.5, 0.7)], offsets=[-5.0, 2.25], seed=12),
]

MAPPING_CASES = [
    dict(name="iso", shape=(18, 14, 22), pix=(1.0, 1.0, 1.0), dim=24, span=22, n_planes=30, C=4,
         views=[(0.3, 0.5, 0.8), (0, 0, 1), (-0.7, 0.1, 0.2)], seed=21),
]


def sampler_volume(case):
    rng = np.random.RandomState(case["seed"])
    vol = rng.randn(*case["shape"], 2).astype(np.float32)
    lab = rng.randint(0, 4, size=case["shape"]).astype(np.uint8)
    affine = np.diag(list(case["pix"]) + [1.0])
    bg = [-1.5, 0.25]
    return vol, lab, affine, bg


def mapping_inputs(case):
    """Per-view prediction stacks [n,dim,dim,C] whose values are exactly representable in float16 (so the
    golden file stores the gather results losslessly in half the bytes)."""
    from oracle import sampler
    rng = np.random.RandomState(case["seed"])
    preds, grids, inv_bases = [], [], []
    hd = case["span"] // 2
    g = np.linspace(-hd, hd, case["dim"])
    for v in case["views"]:
        p = (rng.randint(0, 1024, size=(case["n_planes"], case["dim"], case["dim"], case["C"])) / 1024.0
             ).astype(np.float32)
        preds.append(p)
        grids.append((g, g, sampler.view_offsets(case["dim"], case["span"], case["n_planes"])))
        inv_bases.append(np.linalg.inv(sampler.plane_basis(v)))
    affine = np.diag(list(case["pix"]) + [1.0])
    return preds, grids, inv_bases, case["shape"], affine


# ---- Elastic2D (augmentation/elastic_deformation.py) -------------------------------------------------------
ELASTIC_CASES = [
    dict(name="mild", H=40, W=48, C=2, alpha=40.0, sigma=3.0, bg=[0.25, -1.0], seed=11),
    dict(name="strong", H=64, W=56, C=1, alpha=300.0, sigma=6.5, bg=[0.5], seed=12),   # pushes points out of bounds
]


def elastic_inputs(case):
    rng = np.random.RandomState(100 + case["seed"])
    im = rng.randn(case["H"], case["W"], case["C"]).astype(np.float32)
    lab = rng.randint(0, 5, size=(case["H"], case["W"])).astype(np.uint8)
    return im, lab
