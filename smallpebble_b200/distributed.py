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
    # The steady-state replay launches the all-reduce from CUDA graphs only; NCCL's support for
    # MIXING graph-captured and directly launched collectives on one communicator adds host
    # synchronisation to every launch (measured: update graph 24.8 -> 20.5 us, step 180 -> 174 us
    # on 2 GPUs).  The eager warm-up steps finish (and are synchronised by the first capture)
    # before any graph launches, so the two kinds never overlap.
    os.environ.setdefault("NCCL_GRAPH_MIXING_SUPPORT", "0")
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
    from . import steady

    steady.invalidate()  # recordings made before hold no exchange


def shutdown() -> None:
    import torch.distributed as dist

    _nccl.update(comm=None, tried=False)  # the communicator dies with the process
    _peer.update(tried=False, handle=None, half=0, blocks=None, own=None, plans={})
    _nccl.pop("fast", None)
    if _overlap["side"] is not None:
        _install_hook(False)
    _overlap.update(side=None, mark=None, e_early=None, e_all=None, e_done=None)
    _overlap_reset()
    _plans.clear()
    if _state["owns_group"] and dist.is_initialized():
        dist.destroy_process_group()
    _state.update(initialized=False, rank=0, world=1, backend=None, bucket=None, owns_group=False)
    from . import steady

    steady.invalidate()


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
_nccl = {"comm": None, "tried": False}


def _direct_nccl():
    """The communicator handle of the direct path (an int), or None."""
    if _nccl["tried"]:
        return _nccl["comm"]
    _nccl["tried"] = True
    if os.environ.get("SP_DP_DIRECT_NCCL", "1") == "0" or _state["backend"] != "nccl":
        return None
    import ctypes as C

    import torch
    import torch.distributed as dist

    try:
        # the NCCL torch has loaded (one instance per process); the C ABI binds it by name
        base = os.path.dirname(os.path.dirname(torch.__file__))
        path = os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")
        lib = _abi.load()
        ok = lib.sp_nccl_load(path.encode() if os.path.exists(path) else None) == 0
        uid = (C.c_byte * 128)()
        if ok and _state["rank"] == 0:
            ok = lib.sp_nccl_unique_id(uid) == 0
        payload = [bytes(uid) if ok else None]
        dist.broadcast_object_list(payload, src=0)
        rc = 1
        comm = C.c_void_p()
        if payload[0] is not None and ok:
            C.memmove(uid, payload[0], 128)
            rc = lib.sp_nccl_init(uid, _state["world"], _state["rank"], C.byref(comm))
        # every rank must agree before anyone uses the communicator
        flag = torch.tensor([1 if rc == 0 else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            return None
        _nccl["comm"] = comm.value
        return _nccl["comm"]
    except Exception:  # noqa: BLE001 - any trouble: torch.distributed does the exchange
        _nccl["comm"] = None
        return None


def _allreduce_sum_f32(ptr: int, count: int, tensor_view, stream=None) -> None:
    """In-place sum over ranks of `count` floats at device pointer `ptr`, on the compute stream
    or (direct communicator only) on the raw stream handle `stream`."""
    import torch.distributed as dist

    comm = _direct_nccl()
    if comm is not None:
        _abi.f.sp_allreduce_sum_f32(comm, ptr, count, A.stream() if stream is None else stream)
        return
    dist.all_reduce(tensor_view[:count], op=dist.ReduceOp.SUM)


# --------------------------------------------------------------------------------------------
# Peer-memory exchange: for small gradient buckets (both README models) the all-reduce is not a
# library call at all -- every rank publishes its gradients in an IPC-shared block, and the
# fused Adam kernel of every rank reads all blocks over NVLink and adds them in rank order
# (csrc/optim.cu: adam_multi_peer_kernel).  One kernel does collective + update; nothing is
# latency-bound on a ring or tree.  NCCL stays the path for large buckets, for sgd_step and
# whenever this cannot be set up (SP_DP_PEER=0 turns it off).
# --------------------------------------------------------------------------------------------
_peer = {"tried": False, "handle": None, "half": 0, "blocks": None, "own": None, "plans": {}}
_PEER_MAX_FLOATS = int(os.environ.get("SP_DP_PEER_MAX_FLOATS", str(1 << 20)))


def _peer_exchange(total_floats: int):
    """The exchange handle (an int) if one exists that carries `total_floats`, else None.
    Created collectively on first use (every rank calls this at the same point of the same
    program with the same size)."""
    if total_floats > _PEER_MAX_FLOATS:
        return None
    if _peer["tried"]:
        return _peer["handle"] if total_floats <= _peer["half"] else None
    _peer["tried"] = True
    world = _state["world"]
    if (os.environ.get("SP_DP_PEER", "1") == "0" or _state["backend"] != "nccl" or world < 2
            or world > 8 or int(os.environ.get("LOCAL_WORLD_SIZE", str(world))) != world):
        return None
    import ctypes as C

    import torch
    import torch.distributed as dist

    ok = 1
    own = C.c_void_p()
    handle = (C.c_byte * 64)()
    half = _PEER_MAX_FLOATS
    lib = _abi.load()
    try:
        if lib.sp_peer_alloc(256 + 8 * half, C.byref(own), handle) != 0:
            ok = 0
    except Exception:  # noqa: BLE001
        ok = 0
    gathered = [None] * world
    dist.all_gather_object(gathered, bytes(handle) if ok else None)
    blocks = (C.c_void_p * world)()
    if ok and all(g is not None for g in gathered):
        for r, raw in enumerate(gathered):
            if r == _state["rank"]:
                blocks[r] = own.value
                continue
            h = (C.c_byte * 64).from_buffer_copy(raw)
            p = C.c_void_p()
            if lib.sp_peer_open(h, C.byref(p)) != 0:
                ok = 0
                break
            blocks[r] = p.value
    else:
        ok = 0
    ex = C.c_void_p()
    if ok and lib.sp_peer_exchange_create(world, _state["rank"], half, blocks, C.byref(ex)) != 0:
        ok = 0
    flag = torch.tensor([ok], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # all or nobody (also: nobody starts a step early)
    if int(flag.item()) != 1:
        return None
    _peer.update(handle=ex.value, half=half, blocks=blocks, own=own.value)
    return _peer["handle"] if total_floats <= half else None


def peer_plan(grads):
    """(exchange handle, int64 offsets table) when the gradients of this update travel through
    the peer-memory exchange, else None (the caller all-reduces them with NCCL)."""
    if not (_state["initialized"] and _state["world"] > 1) or len(grads) > 24:
        return None
    shapes = tuple([g.shape for g in grads])
    plan = _peer["plans"].get(shapes)
    if plan is None:
        layout = bucket_layout([g.size for g in grads])
        handle = _peer_exchange(layout[-1])
        plan = _peer["plans"][shapes] = (handle, _abi.dims(layout[:-1])) if handle else False
    return plan or None


def exchange_description() -> str:
    if _peer["handle"] is not None:
        return ("one-shot all-reduce over NVLink peer memory fused into the Adam kernel "
                "(buckets <= %d floats); NCCL all-reduce otherwise" % _peer["half"])
    return "NCCL all-reduce of one flat gradient bucket"


def bucket_layout(sizes) -> list[int]:
    """Element offset of each gradient inside the flat bucket (16-byte aligned slots so the
    fused Adam kernel keeps its 128-bit accesses)."""
    offsets, cur = [], 0
    for n in sizes:
        offsets.append(cur)
        cur += (int(n) + 3) // 4 * 4
    return offsets + [cur]


_plans = {}  # gradient shapes -> bucket plan
_retired_buckets = []


def _bucket_plan(shapes, sizes):
    plan = _plans.get(shapes)
    if plan is None:
        layout = bucket_layout(sizes)
        plan = _plans[shapes] = {
            "layout": layout, "total": layout[-1], "c_sizes": _abi.dims(sizes),
            "c_offsets": _abi.dims(layout[:-1]), "ptrs": None, "table": None, "views": None,
            "base": None, "split": None}
    return plan


# --------------------------------------------------------------------------------------------
# Overlap of the exchange with the backward pass.  The collective is latency-bound (29 us for
# the README CNN's 0.79 MB bucket on 8 GPUs) and at N = 8 the step is device-bound by exactly
# that, while most of the bucket -- the classifier and the last conv layer's filters, the
# first gradients the sweep produces -- is ready long before the backward pass ends.  So the
# bucket is split in two by PRODUCTION order (early part: everything but the last <= 10 % of
# the bytes): core.get_gradients reports every write to a
# learnable's gradient (core._dp_hook); after the write that completes the early part an event
# is recorded, and the side stream packs + all-reduces that part while the main stream is
# still computing the remaining gradients; the late part follows on the same side stream (one
# stream for all NCCL calls: same order on every rank), and the optimiser waits for both.
# The split is planned from the first step's observed order and re-validated every step
# (every early gradient must be the very array the sweep last wrote before the event);
# anything unexpected -- a different tape, gradients built outside the sweep, CUDA-graph
# capture -- takes the single all-reduce on the compute stream.
# --------------------------------------------------------------------------------------------
# Off unless SP_DP_OVERLAP=1 ("auto": buckets of >= _OVERLAP_MIN_FLOATS only).  Measured on two
# GPUs it does not pay: the overlapped path costs ~25 us more host time per step (write hook,
# two packs, two NCCL calls, four event operations), which the host-bound README CNN cannot
# afford (0.177 -> 0.192 ms per step), and the device-bound VGG-6 step did not change within
# noise (2.423 vs 2.433 ms: its 60 us of data-parallel cost is not the wait for the collective).
# On 8 GPUs, where the README CNN's step IS device-bound by the 29 us collective, two
# latency-bound collectives cost more than the one they replace: 0.198 -> 0.232 ms per step.
# What would pay there is a cheaper collective (one-shot reduction over NVLink peer memory),
# not a split one.
_OVERLAP_MIN_FLOATS = 1 << 19
_overlap = {"enabled": os.environ.get("SP_DP_OVERLAP", "0"), "seq": 0, "written": {},
            "mark": None, "event_seq": -1, "e_early": None, "e_all": None, "e_done": None,
            "side": None, "used": 0}


def _on_gradient_written(var, value) -> None:
    """core.get_gradients calls this after every write to a learnable leaf's gradient."""
    o = _overlap
    o["seq"] = seq = o["seq"] + 1
    o["written"][var] = (seq, id(value))
    if var is o["mark"]:
        value.ptr  # a deferred gradient is launched now, before the event
        _abi.f.sp_event_record(o["e_early"], A.stream())
        o["event_seq"] = seq


def _overlap_reset() -> None:
    o = _overlap
    o["seq"] = 0
    o["written"].clear()
    o["event_seq"] = -1


def _install_hook(enable: bool) -> None:
    from . import core

    core._dp_hook = _on_gradient_written if enable else None


def _plan_split(plan, grads, variables):
    """Split the bucket by production order: (order, k) = gradient indices sorted by the
    sequence number of their last write, the first k of them forming the early part.  None
    when the order is unknown."""
    written = _overlap["written"]
    n = len(grads)
    if n < 2 or any(v not in written for v in variables):
        return None
    order = sorted(range(n), key=lambda i: written[variables[i]][0])
    total = sum(g.size for g in grads)
    # the late part: the last-produced gradients, as few bytes as possible (its all-reduce is
    # the exposed one) -- at most 10 % of the bucket, but at least the very last gradient
    k, late = n - 1, grads[order[n - 1]].size
    while k > 1 and (late + grads[order[k - 1]].size) * 10 <= total:
        k -= 1
        late += grads[order[k]].size
    sizes = [grads[i].size for i in order]
    layout = bucket_layout(sizes)
    offsets = [0] * n
    for pos, i in enumerate(order):
        offsets[i] = layout[pos]
    return {
        "order": order, "k": k, "offsets": offsets, "total": layout[-1],
        "early_count": layout[k], "late_off": layout[k], "late_count": layout[-1] - layout[k],
        "c_sizes_a": _abi.dims(sizes[:k]), "c_off_a": _abi.dims(layout[:k]),
        "c_sizes_b": _abi.dims(sizes[k:]), "c_off_b": _abi.dims(layout[k:-1]),
        "mark": variables[order[k - 1]], "ptrs": None, "tab_a": None, "tab_b": None,
        "views": None, "base": None,
    }


def _allreduce_overlapped(plan, split, grads, variables, bucket):
    """Early part on the side stream behind the marker event, late part behind the end of the
    backward pass; the compute stream waits for both.  Returns the bucket views or None when
    this step does not match the plan (the caller then uses the single all-reduce)."""
    import ctypes as C

    o = _overlap
    written = o["written"]
    k, order = split["k"], split["order"]
    if o["event_seq"] <= 0:
        return None
    for pos in range(k):
        i = order[pos]
        rec = written.get(variables[i])
        if rec is None or rec[0] > o["event_seq"] or rec[1] != id(grads[i]):
            return None
    side_raw = o["side"]
    base = bucket.data_ptr()
    ptrs = tuple([grads[i].ptr for i in order])
    if ptrs != split["ptrs"]:
        split["ptrs"] = ptrs
        split["tab_a"] = (C.c_void_p * k)(*ptrs[:k])
        split["tab_b"] = (C.c_void_p * (len(ptrs) - k))(*ptrs[k:])
    F = _abi.f
    main = A.stream()
    F.sp_event_record(o["e_all"], main)  # every gradient kernel has been enqueued before this
    F.sp_stream_wait_event(side_raw, o["e_early"])
    F.sp_pack_multi(k, split["tab_a"], split["c_sizes_a"], split["c_off_a"], base, side_raw)
    _allreduce_sum_f32(base, split["early_count"], bucket, side_raw)
    F.sp_stream_wait_event(side_raw, o["e_all"])
    F.sp_pack_multi(len(ptrs) - k, split["tab_b"], split["c_sizes_b"], split["c_off_b"], base,
                    side_raw)
    _allreduce_sum_f32(base + 4 * split["late_off"], split["late_count"], bucket, side_raw)
    F.sp_event_record(o["e_done"], side_raw)
    F.sp_stream_wait_event(main, o["e_done"])
    o["used"] += 1
    if split["base"] != base:
        split["base"] = base
        split["views"] = [DeviceArray(g.shape, A.FLOAT, bucket, base + 4 * off)
                          for g, off in zip(grads, split["offsets"])]
    return split["views"]


def allreduce_gradients(grads, variables=None):
    """Sum each gradient over all ranks.  Returns arrays that alias the flat bucket.
    `variables` (the learnables the gradients belong to, same order) lets the exchange overlap
    the backward pass (see above); without it the whole bucket is reduced in one call."""
    import torch
    import torch.distributed as dist

    sizes = [g.size for g in grads]
    on_gpu = _state["backend"] == "nccl"
    if on_gpu:
        # everything that is the same every step is built once per set of gradient shapes: the
        # bucket layout, the size / offset tables of the pack kernel, its pointer table (the
        # gradient buffers come out of the recycling pool in the same order every step) and
        # the views of the bucket that are handed to the optimiser
        plan = _bucket_plan(tuple([g.shape for g in grads]), sizes)
        total = plan["total"]
        bucket = _state["bucket"]
        if bucket is None or bucket.numel() < total + 64:
            if bucket is not None:
                _retired_buckets.append(bucket)  # recorded CUDA graphs may still write into it
            bucket = torch.empty(max(total, 1) + 64, dtype=torch.float32, device=A.require_cuda())
            _state["bucket"] = bucket
        base = bucket.data_ptr()
        o = _overlap
        wanted = o["enabled"] == "1" or (o["enabled"] == "auto" and total >= _OVERLAP_MIN_FLOATS)
        if (wanted and variables is not None and world_size() > 1 and not A._pool_bypass
                and _direct_nccl() is not None):
            if o["side"] is None:
                import ctypes as C

                slot = (C.c_void_p * 1)()  # an array: both bindings pass its address
                _abi.f.sp_stream_create(slot)
                o["side"] = slot[0]
                for name in ("e_early", "e_all", "e_done"):
                    _abi.f.sp_event_create(slot)
                    o[name] = slot[0]
                _install_hook(True)
            split = plan["split"]
            views = None
            if split is not None and split["mark"] is o["mark"]:
                views = _allreduce_overlapped(plan, split, grads, variables, bucket)
            elif o["written"]:
                # first step with this set of gradients (or another model took the marker):
                # plan the split from the order the sweep just produced them in
                split = plan["split"] = _plan_split(plan, grads, variables)
                o["mark"] = split["mark"] if split is not None else None
            _overlap_reset()
            if views is not None:
                return views
        ptrs = tuple([g.ptr for g in grads])
        if ptrs != plan["ptrs"]:
            import ctypes as C

            plan["ptrs"], plan["table"] = ptrs, (C.c_void_p * len(ptrs))(*ptrs)
        _abi.f.sp_pack_multi(len(grads), plan["table"], plan["c_sizes"], plan["c_offsets"], base,
                             A.stream())
        _allreduce_sum_f32(base, total, bucket)
        if plan["base"] != base:
            plan["base"] = base
            plan["views"] = [DeviceArray(g.shape, A.FLOAT, bucket, base + 4 * off)
                             for g, off in zip(grads, plan["layout"])]
        return plan["views"]
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
