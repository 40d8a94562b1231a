"""Exact centre of the real-space voxel grid, as the reference computes it.

`get_voxel_grid_real_space` (mpunet/interpolation/sample_grid.py:101-130) centres the grid with
`np.mean(grid_points_real_space, axis=0)` over all prod(shape) voxels.  The closed form A.(n-1)/2 is the same number
mathematically but not bit for bit (numpy's pairwise summation rounds differently: 1e-15 off for a rotated affine),
and a nearest-neighbour lookup decides ties on exactly those bits.  This module reproduces numpy's result exactly:
the recursion tree of `pairwise_sum` is derived here from the element count, its <= 128-element leaves are summed on
the device in numpy's operation order (mpu_voxel_leaf_sums), and the tree is combined on the host.
"""
import ctypes
import functools

import numpy as np

PW_BLOCKSIZE = 128


@functools.lru_cache(maxsize=16)
def pairwise_tree(n):
    """Recursion tree of numpy's pairwise_sum over n elements.
    Returns (leaf_start int64[L], leaf_len int32[L], levels) where levels is a list (root first) of
    (node_ids, left_ids, right_ids) for the inner nodes created at that depth, node ids index a flat array in which
    leaf t has id `leaf_id[t]` (returned as 4th element) ."""
    starts = [np.array([0], dtype=np.int64)]
    lens = [np.array([n], dtype=np.int64)]
    ids = [np.array([0], dtype=np.int64)]
    next_id = 1
    levels = []
    leaf_start, leaf_len, leaf_id = [], [], []
    while len(starts[-1]):
        s, l, i = starts[-1], lens[-1], ids[-1]
        is_leaf = l <= PW_BLOCKSIZE
        leaf_start.append(s[is_leaf])
        leaf_len.append(l[is_leaf])
        leaf_id.append(i[is_leaf])
        s, l, i = s[~is_leaf], l[~is_leaf], i[~is_leaf]
        n2 = l // 2
        n2 = n2 - n2 % 8
        left = next_id + 2 * np.arange(len(s), dtype=np.int64)
        right = left + 1
        next_id += 2 * len(s)
        levels.append((i, left, right))
        starts.append(np.stack([s, s + n2], 1).ravel())
        lens.append(np.stack([n2, l - n2], 1).ravel())
        ids.append(np.stack([left, right], 1).ravel())
    return (np.concatenate(leaf_start), np.concatenate(leaf_len).astype(np.int32), levels, np.concatenate(leaf_id),
            next_id)


def combine(n, leaf_sums):
    """Total of numpy's pairwise sum over n elements given the leaf sums (ordered as pairwise_tree's leaves)."""
    _, _, levels, leaf_id, n_nodes = pairwise_tree(int(n))
    val = np.zeros(n_nodes, dtype=np.float64)
    val[leaf_id] = leaf_sums
    for node, left, right in reversed(levels):
        val[node] = val[left] + val[right]
    return val[0]


def voxel_grid_center_exact(shape3, affine3x3, device=None):
    """np.mean(A . ijk, axis=0) over the whole voxel grid, bit for bit (float64 [3])."""
    import torch
    from .. import _C
    from .._C import lib, check
    shape3 = tuple(int(v) for v in shape3)
    n = int(np.prod(shape3))
    ls, ll, _, _, _ = pairwise_tree(n)
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    ls_d = torch.from_numpy(ls).to(dev)
    ll_d = torch.from_numpy(ll).to(dev)
    out = torch.empty(3 * len(ls), dtype=torch.float64, device=dev)
    A = np.ascontiguousarray(np.asarray(affine3x3, dtype=np.float64)[:3, :3])
    check(lib.mpu_voxel_leaf_sums(_C.int_array(shape3), _C.double_array(A.ravel()), _C.ptr(ls_d), _C.ptr(ll_d),
                                  ctypes.c_longlong(len(ls)), _C.ptr(out), _C.current_stream()), "mpu_voxel_leaf_sums")
    sums = out.cpu().numpy().reshape(3, len(ls))
    return np.array([combine(n, sums[r]) / n for r in range(3)])
