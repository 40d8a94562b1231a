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
    # rotated + sheared + anisotropic affine: real-space coordinates and the grid centre are no longer exact in binary
    dict(name="rot", shape=(17, 15, 21), pix=None, dim=26, span=28, n_planes=33, C=5,
         views=[(0.3, 0.5, 0.8), (0, 0, 1), (0.6, -0.2, 0.4)], seed=22),
]


def rotated_affine():
    th = 0.3
    R = np.array([[np.cos(th), -np.sin(th), 0.1], [np.sin(th), np.cos(th), 0], [0.02, 0.05, 1]])
    A = np.eye(4)
    A[:3, :3] = R.dot(np.diag([1.0, 0.7, 1.3]))
    return A


def rotated_affine_dyadic():
    """A genuinely rotated affine (3-4-5 rotations about z and x) whose column norms come out EXACTLY (1, 0.5, 2):
    the reference multiplies its float32 voxel axes by these float64 pixdims (sample_grid.py:83-85), which stays
    float32 under the numpy 1.x it pins and becomes float64 under numpy >= 2 (NEP 50) - identical values only when
    the products are exact.  The view-stack golden uses this affine so that it pins the reference's behaviour in
    both numpy generations (the product follows the pinned 1.x semantics: float32 axes)."""
    Rz = np.array([[0.6, -0.8, 0], [0.8, 0.6, 0], [0, 0, 1.0]])
    Rx = np.array([[1.0, 0, 0], [0, 0.8, -0.6], [0, 0.6, 0.8]])
    A = np.eye(4)
    A[:3, :3] = Rz.dot(Rx).dot(np.diag([1.0, 0.5, 2.0]))
    return A


VOXEL_CENTER_CASES = [((18, 14, 22), "diag"), ((17, 15, 21), "rot"), ((64, 64, 64), "rot"), ((37, 129, 5), "rot"),
                      ((100, 90, 80), "rot"), ((7, 3, 2), "rot")]


def voxel_center_affine(kind):
    return rotated_affine() if kind == "rot" else np.diag([1.0, 0.7, 2.0, 1.0])


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
    affine = rotated_affine() if case["pix"] is None else np.diag(list(case["pix"]) + [1.0])
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


# ---- inference view stacks (sequences/isotrophic_live_view_sequence_2d.py:29-117 + preprocessing/scaling.py) ----------
VIEW_STACK_CASES = [
    dict(name="iso", shape=(18, 14, 22), affine="diag", dim=24, span=22, n_planes="same+20", view=(0.3, 0.5, 0.8), seed=31),
    dict(name="rot", shape=(16, 20, 12), affine="rot", dim=20, span=24, n_planes=9, view=(0.05, 0.1, 0.99), seed=32),
]


def view_stack_inputs(case):
    rng = np.random.RandomState(case["seed"])
    vol = (rng.randn(*case["shape"], 2) * [3.0, 0.5] + [10.0, -1.0]).astype(np.float32)
    lab = rng.randint(0, 4, size=case["shape"]).astype(np.uint8)
    affine = rotated_affine_dyadic() if case["affine"] == "rot" else np.diag([1.0, 1.0, 1.0, 1.0])
    bg = [float(np.float32(np.percentile(vol[..., c], 1))) for c in range(2)]
    return vol, lab, affine, bg


# ---- training-batch rejection rules (sequences/isotrophic_live_view_sequence_2d.py:119-161 and
#      isotrophic_live_view_sequence.py:91-128) driven from a fixed candidate list --------------------------------------
BATCH_RULE_CASES = [
    dict(name="sparse_fg", shape=(24, 24, 24), dim=24, span=40, B=12, n_classes=5, tries=10, fg_frac=0.5, seed=41),
    dict(name="dense_small_batch", shape=(20, 20, 20), dim=16, span=20, B=3, n_classes=5, tries=10, fg_frac=0.7, seed=42),
]


def batch_rule_inputs(case):
    """Volume whose foreground classes are small blobs (so that many candidate planes miss them), a span larger
    than the volume (so that some planes are entirely out of bounds: is_valid_im), and the candidate list
    [B][tries] = (view index, offset, normal noise[3]) every implementation is driven from."""
    rng = np.random.RandomState(case["seed"])
    X, Y, Z = case["shape"]
    vol = rng.randn(X, Y, Z, 1).astype(np.float32)
    lab = np.zeros(case["shape"], np.uint8)
    for c in range(1, case["n_classes"]):
        ctr = rng.randint(4, X - 4, 3)
        r = 2 + (c % 2)
        lab[ctr[0] - r:ctr[0] + r, ctr[1] - r:ctr[1] + r, ctr[2] - r:ctr[2] + r] = c
    views = np.array([[0.3, 0.5, 0.8], [0, 0, 1.0], [1.0, 0, 0], [-0.7, 0.1, 0.2]])
    views = views / np.linalg.norm(views, axis=1, keepdims=True)
    sphere_r = case["span"] // 2
    cand_view = rng.randint(0, len(views), size=(case["B"], case["tries"]))
    cand_off = rng.uniform(-sphere_r, sphere_r, size=(case["B"], case["tries"]))
    cand_noise = rng.normal(scale=0.1, size=(case["B"], case["tries"], 3))
    bg = [float(np.float32(np.percentile(vol[..., 0], 1)))]
    return vol, lab, views, cand_view, cand_off, cand_noise, bg


# ---- U-Net graph fixtures (tests/golden/unet_graph_*.npz, oracle/make_golden.py::unet_graph_goldens) ----
UNET_GRAPH_CASES = {
    "small": dict(n_classes=3, dim=32, n_channels=1, depth=4, complexity_factor=0.125),
    "rgb": dict(n_classes=2, dim=48, n_channels=3, depth=4, complexity_factor=0.25),
    "benchmark": dict(n_classes=5, dim=256, n_channels=1, depth=4, complexity_factor=2.0),
}


def unet_graph_input(kw, batch=2):
    """Seeded input batch [B, dim, dim, n_channels] float32 of a U-Net graph case."""
    rng = np.random.RandomState(7 + kw["dim"])
    return rng.randn(batch, kw["dim"], kw["dim"], kw["n_channels"]).astype(np.float32)


# ---- fusion layer / generalized dice loss fixtures (tests/golden/fusion_ref.npz) ----
def fusion_inputs(n=4096, n_views=6, n_classes=5, seed=11):
    """Seeded softmax-like view predictions x [n, V, C] f32, labels y [n, 1] u8, weights W [V, C], bias b [1, C]."""
    rng = np.random.RandomState(seed)
    x = rng.dirichlet(0.4 * np.ones(n_classes), size=(n, n_views)).astype(np.float32)
    y = rng.randint(0, n_classes, size=(n, 1)).astype(np.uint8)
    W = rng.uniform(0.5, 1.5, (n_views, n_classes)).astype(np.float32)
    b = (0.1 * rng.randn(1, n_classes)).astype(np.float32)
    return x, y, W, b
