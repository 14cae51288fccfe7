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

    _nccl.update(comm=None, tried=False)  # the communicator dies with the process
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


# --------------------------------------------------------------------------------------------
# Direct NCCL: the gradient all-reduce is issued straight on OUR compute stream through the NCCL
# library torch has already loaded (ctypes, same .so instance), with a communicator of its own
# bootstrapped over torch.distributed.  `dist.all_reduce` costs ~35 us of host time per call
# and hops through ProcessGroupNCCL's internal stream with two event waits -- a large part of
# the ~60 us the exchange added to a 0.25 ms step.  torch.distributed stays the bootstrap and
# the fallback (SP_DP_DIRECT_NCCL=0, or any failure while setting this up).
# --------------------------------------------------------------------------------------------
_nccl = {"lib": None, "comm": None, "tried": False}


def _direct_nccl():
    """(lib, comm) of the direct communicator, or None."""
    if _nccl["tried"]:
        return (_nccl["lib"], _nccl["comm"]) if _nccl["comm"] is not None else None
    _nccl["tried"] = True
    if os.environ.get("SP_DP_DIRECT_NCCL", "1") == "0" or _state["backend"] != "nccl":
        return None
    import ctypes as C

    import torch
    import torch.distributed as dist

    try:
        lib = None
        base = os.path.dirname(os.path.dirname(torch.__file__))
        for cand in (os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2"), "libnccl.so.2"):
            try:
                lib = C.CDLL(cand)
                break
            except OSError:
                continue
        if lib is None:
            return None

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_byte * 128)]

        lib.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
        lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        lib.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p]
        for fn in (lib.ncclGetUniqueId, lib.ncclCommInitRank, lib.ncclAllReduce):
            fn.restype = C.c_int
        uid = UniqueId()
        if _state["rank"] == 0 and lib.ncclGetUniqueId(C.byref(uid)) != 0:
            ok = False
        else:
            ok = True
        payload = [bytes(uid.internal) if ok else None]
        dist.broadcast_object_list(payload, src=0)
        if payload[0] is None:
            return None
        C.memmove(C.byref(uid), payload[0], 128)
        comm = C.c_void_p()
        rc = lib.ncclCommInitRank(C.byref(comm), _state["world"], uid, _state["rank"])
        # every rank must agree before anyone uses the communicator
        flag = torch.tensor([1 if rc == 0 else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            return None
        _nccl["lib"], _nccl["comm"] = lib, comm
        # the same vectorcall trampoline the C ABI uses (7 integer-class arguments): ~0.3 us
        # per call instead of ~4 us of ctypes argument conversion
        fast = _abi._load_fast()
        if fast is not None:
            def _nccl_failed(name, rc):
                raise RuntimeError(f"{name} failed with code {rc}")

            _nccl["fast"] = fast.bind(C.cast(lib.ncclAllReduce, C.c_void_p).value, "ppppppp",
                                      "ncclAllReduce", _nccl_failed, False)
        return lib, comm
    except Exception:  # noqa: BLE001 - any trouble: torch.distributed does the exchange
        _nccl["comm"] = None
        return None


def _allreduce_sum_f32(ptr: int, count: int, tensor_view) -> None:
    """In-place sum over ranks of `count` floats at device pointer `ptr` on the compute stream."""
    import torch.distributed as dist

    direct = _direct_nccl()
    if direct is not None:
        lib, comm = direct
        call = _nccl.get("fast")
        if call is not None:
            call(ptr, ptr, count, 7, 0, comm.value, A.stream())  # ncclFloat32, ncclSum
            return
        rc = lib.ncclAllReduce(ptr, ptr, count, 7, 0, comm, A.stream())
        if rc != 0:
            raise RuntimeError(f"ncclAllReduce failed with code {rc}")
        return
    dist.all_reduce(tensor_view[:count], op=dist.ReduceOp.SUM)


def bucket_layout(sizes) -> list[int]:
    """Element offset of each gradient inside the flat bucket (16-byte aligned slots so the
    fused Adam kernel keeps its 128-bit accesses)."""
    offsets, cur = [], 0
    for n in sizes:
        offsets.append(cur)
        cur += (int(n) + 3) // 4 * 4
    return offsets + [cur]


_plans = {}  # gradient shapes -> (layout, total, ctypes size / offset tables)


def _bucket_plan(shapes, sizes):
    plan = _plans.get(shapes)
    if plan is None:
        layout = bucket_layout(sizes)
        plan = _plans[shapes] = (layout, layout[-1], _abi.dims(sizes), _abi.dims(layout[:-1]),
                                 {"ptrs": None, "table": None, "views": None, "base": None})
    return plan


def allreduce_gradients(grads):
    """Sum each gradient over all ranks.  Returns arrays that alias the flat bucket."""
    import torch
    import torch.distributed as dist

    sizes = [g.size for g in grads]
    on_gpu = _state["backend"] == "nccl"
    if on_gpu:
        # everything that is the same every step is built once per set of gradient shapes: the
        # bucket layout, the size / offset tables of the pack kernel, its pointer table (the
        # gradient buffers come out of the recycling pool in the same order every step) and
        # the views of the bucket that are handed to the optimiser
        layout, total, c_sizes, c_offsets, cache = _bucket_plan(tuple([g.shape for g in grads]),
                                                               sizes)
        bucket = _state["bucket"]
        if bucket is None or bucket.numel() < total:
            bucket = torch.empty(max(total, 1), dtype=torch.float32, device=A.require_cuda())
            _state["bucket"] = bucket
        base = bucket.data_ptr()
        ptrs = tuple([g.ptr for g in grads])
        if ptrs != cache["ptrs"]:
            import ctypes as C

            cache["ptrs"], cache["table"] = ptrs, (C.c_void_p * len(ptrs))(*ptrs)
        _abi.f.sp_pack_multi(len(grads), cache["table"], c_sizes, c_offsets, base, A.stream())
        _allreduce_sum_f32(base, total, bucket)
        if cache["base"] != base:
            cache["base"] = base
            cache["views"] = [DeviceArray(g.shape, A.FLOAT, bucket, base + 4 * off)
                              for g, off in zip(grads, layout)]
        return cache["views"]
    layout = bucket_layout(sizes)
    total = layout[-1]
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
