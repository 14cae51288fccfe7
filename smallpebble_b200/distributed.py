"""Batch-sharded data parallelism: one process per GPU, `torch.distributed` (NCCL over
NVLink 5 / NVSwitch) as plumbing, ONE all-reduce of one flat gradient bucket per step.

The path shards over the batch dimension and has exactly one exchange step: the loss is a
SUM over the batch (sp.py:621), so grad(global batch) = sum over ranks of grad(shard) --
all-reduce with SUM and no rescaling reproduces the single-GPU step on the global batch.
The reference has no counterpart (it is single-process NumPy).

Usage (the training script itself is unchanged SmallPebble):

    import smallpebble_b200 as sp
    sp.distributed.init()                       # reads RANK / LOCAL_RANK / WORLD_SIZE
    xs, ys = sp.distributed.shard(xbatch, ybatch)   # this rank's contiguous slice
    ...
    adam.training_step(learnables, gradients)   # all-reduces, then the fused Adam kernel

On a machine without CUDA `init()` selects gloo and the bucket lives in host memory: that is
how the plumbing (bucket layout, shard arithmetic, reduction) is tested at world_size 2.
"""

from __future__ import annotations

import math
import os

import numpy as np

from . import _abi
from . import array as A
from .array import DeviceArray

_state = {"initialized": False, "rank": 0, "world": 1, "backend": None, "bucket": None,
          "owns_group": False}


def is_initialized() -> bool:
    return _state["initialized"]


def rank() -> int:
    return _state["rank"]


def world_size() -> int:
    return _state["world"]


def init(backend: str | None = None) -> None:
    """Join the process group described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rk)))
    use_cuda = torch.cuda.is_available()
    if backend is None:
        backend = "nccl" if use_cuda else "gloo"
    if use_cuda:
        torch.cuda.set_device(local % torch.cuda.device_count())
        A.reset_device_cache()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group(backend=backend, rank=rk, world_size=world, **kwargs)
        _state["owns_group"] = True
    _state.update(initialized=True, rank=rk, world=world, backend=backend, bucket=None)


def shutdown() -> None:
    import torch.distributed as dist

    if _state["owns_group"] and dist.is_initialized():
        dist.destroy_process_group()
    _state.update(initialized=False, rank=0, world=1, backend=None, bucket=None, owns_group=False)


def shard_bounds(n: int, rk: int | None = None, world: int | None = None) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of a global batch of n owned by rank rk: the first n % world
    ranks get one extra row, so every row is owned exactly once."""
    rk = rank() if rk is None else rk
    world = world_size() if world is None else world
    base, extra = divmod(n, world)
    lo = rk * base + min(rk, extra)
    return lo, lo + base + (1 if rk < extra else 0)


def shard(*arrays):
    """This rank's rows of each global-batch array."""
    out = []
    for a in arrays:
        lo, hi = shard_bounds(len(a))
        out.append(a[lo:hi])
    return out[0] if len(out) == 1 else tuple(out)


def bucket_layout(sizes) -> list[int]:
    """Element offset of each gradient inside the flat bucket (16-byte aligned slots so the
    fused Adam kernel keeps its 128-bit accesses)."""
    offsets, cur = [], 0
    for n in sizes:
        offsets.append(cur)
        cur += (int(n) + 3) // 4 * 4
    return offsets + [cur]


def allreduce_gradients(grads):
    """Sum each gradient over all ranks.  Returns arrays that alias the flat bucket."""
    import torch
    import torch.distributed as dist

    sizes = [g.size for g in grads]
    layout = bucket_layout(sizes)
    total = layout[-1]
    on_gpu = _state["backend"] == "nccl"
    if on_gpu:
        bucket = _state["bucket"]
        if bucket is None or bucket.numel() < total:
            bucket = torch.empty(max(total, 1), dtype=torch.float32, device=A.require_cuda())
            _state["bucket"] = bucket
        base = bucket.data_ptr()
        import ctypes as C

        _abi.f.sp_pack_multi(len(grads), (C.c_void_p * len(grads))(*[g.ptr for g in grads]),
                             _abi.dims(sizes), _abi.dims(layout[:-1]), base, A.stream())
        dist.all_reduce(bucket[:total], op=dist.ReduceOp.SUM)
        return [DeviceArray(g.shape, A.FLOAT, bucket, base + 4 * off)
                for g, off in zip(grads, layout)]
    # host path (gloo): plumbing tests on machines without a GPU
    bucket = torch.zeros(max(total, 1), dtype=torch.float32)
    view = bucket.numpy()
    for g, off in zip(grads, layout):
        view[off:off + g.size] = np.asarray(g, dtype=np.float32).reshape(-1)
    if world_size() > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
    return [DeviceArray.from_host(view[off:off + g.size].reshape(g.shape))
            for g, off in zip(grads, layout)]


def allreduce_scalar(value: float) -> float:
    """Sum of a host scalar over ranks (e.g. the global-batch loss for logging)."""
    import torch
    import torch.distributed as dist

    if world_size() == 1:
        return float(value)
    dev = A.require_cuda() if _state["backend"] == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def broadcast_parameters(variables, src: int = 0) -> None:
    """Make every rank start from rank `src`'s weights (not needed when all ranks build the
    model after the same np.random.seed)."""
    import torch
    import torch.distributed as dist

    if world_size() == 1:
        return
    on_gpu = _state["backend"] == "nccl"
    for v in variables:
        arr = v.array
        if on_gpu:
            t = torch.empty(max(arr.size, 1), dtype=torch.float32, device=A.require_cuda())
            _abi.f.sp_memcpy_d2d(t.data_ptr(), arr.ptr, arr.nbytes, A.stream())
            dist.broadcast(t, src=src)
            _abi.f.sp_memcpy_d2d(arr.ptr, t.data_ptr(), arr.nbytes, A.stream())
        else:
            t = torch.from_numpy(np.asarray(arr, dtype=np.float32).reshape(-1).copy())
            dist.broadcast(t, src=src)
            v.array = t.numpy().reshape(arr.shape)


def barrier() -> None:
    import torch.distributed as dist

    if world_size() > 1:
        dist.barrier()
