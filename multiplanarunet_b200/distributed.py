"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch on the GPU
box, gloo in CPU tests).  Replaces tf.distribute.MirroredStrategy as the reference uses it
(mpunet/bin/train.py:349-358, bin/predict.py:214, bin/train_fusion.py:336): the path shards by
VOLUME (one resident volume per rank, no data-path collective) and has exactly one exchange step per
train step - the SUM all-reduce of the flat fp32 gradient buffer (BatchNorm statistics stay per
replica, as MirroredStrategy's non-synchronised BatchNormalization does).
"""
import os


def is_initialized():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def init_from_env(backend=None, device=None):
    """Initialise the default process group from torchrun's environment (no-op for world size 1)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or is_initialized():
        return world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl" and device is not None:
        kw["device_id"] = device
    dist.init_process_group(backend, **kw)
    return world


def world_size():
    import torch.distributed as dist
    return dist.get_world_size() if is_initialized() else 1


def rank():
    import torch.distributed as dist
    return dist.get_rank() if is_initialized() else 0


def shard(items, world=None, rnk=None):
    """Round-robin shard of independent units (volumes, views, point sets) across ranks."""
    world = world_size() if world is None else world
    rnk = rank() if rnk is None else rnk
    return [it for i, it in enumerate(items) if i % world == rnk]


def all_reduce_flat(flat, bucket_elems=8 * 1024 * 1024, average=False):
    """SUM all-reduce of a flat parameter-gradient tensor in buckets (32 MB of fp32 by default) issued
    asynchronously so NCCL pipelines them; returns after all buckets completed."""
    import torch.distributed as dist
    if not is_initialized() or dist.get_world_size() == 1:
        return flat
    handles = []
    n = flat.numel()
    for s in range(0, n, bucket_elems):
        handles.append(dist.all_reduce(flat[s:min(n, s + bucket_elems)], async_op=True))
    for h in handles:
        h.wait()
    if average:
        flat.div_(dist.get_world_size())
    return flat


def broadcast_flat(flat, src=0):
    import torch.distributed as dist
    if is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src)
    return flat


class DataParallel(object):
    """Wraps a multiplanarunet_b200.models.UNet: identical initial weights on every rank, gradients
    summed across ranks before the (identical) Adam update - weights stay bit-identical."""

    def __init__(self, model):
        self.model = model
        broadcast_flat(model.params)
        broadcast_flat(model.bn_state)
        model._sync()

    def train_on_batch(self, x, y, sample_weight=None):
        # gradients are SUM-reduced range by range while backward is still running
        loss = self.model.forward_backward_overlapped(x, y, sample_weight)
        self.model.apply_gradients()
        H, W, _ = self.model.img_shape
        return float(loss.item()) / (self.model._last_B * H * W)
