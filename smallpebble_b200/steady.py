"""Steady-state replay of the unchanged training loop (VERDICT r01 #2c).

The README loop (README.md:310-325) is three API calls per step --

    loss_val = loss.run();  gradients = sp.get_gradients(loss_val);  adam.training_step(learnables, gradients)

-- that issue the same ~20 kernels with the same shapes every iteration; on a B200 the Python
and launch cost of issuing them (~160 us) is longer than the kernels take.  Once a root lazy
node has been run a few times with unchanged input shapes, each of the three calls is recorded
into a CUDA graph of its own (forward F, backward B, update U) and later calls REPLAY them:

  * `root.run()`      copies the placeholder values into static input buffers, replays F and
                      returns a fresh Variable holding a copy of the root value; its tape is
                      the recorded one (static buffers);
  * `get_gradients()` of that Variable replays B and returns a dict over the same buffers (the
                      gradient of a non-learnable leaf stays deferred, as in eager mode);
  * `Adam.training_step()` given that dict replays U.

Nothing is fused or skipped: the graphs hold exactly the kernels the eager calls launch, so
the results are bit-identical to eager mode.  Two things are done that the eager loop cannot
afford on the host: in B the parameter-gradient kernels (conv2d wgrad, the matmul weight
gradient, bias sums) are forked onto a second stream and run beside the data-gradient chain
they do not feed (joined before B ends), and all launch overhead disappears.

Reference semantics that are kept: Adam REBINDS `variable.array` every step (sp.py:583) -- U
writes the new parameters into the arrays of two steps ago (two recordings per root, keyed by
the parameter buffers they read), so whatever still reads the step's own weights after the
update (the deferred leaf gradient) sees the right values.  What changes: arrays reachable
from a step's results, other than the root value itself, are overwritten by the next `run()`
of the same root (hold `np.asarray(...)` copies instead), and a parameter array object comes
round again every second step.  Any deviation from the recorded pattern -- new shapes, a
rebound parameter, another optimiser, changed precision / fusion settings -- falls back to the
eager path for that call and re-records when the steps have settled again.
`sp.set_steady_replay(False)` or SP_STEADY=0 turns the whole mechanism off.
"""

from __future__ import annotations

import ctypes as C
import gc
import os

import numpy as np

from . import _abi
from . import array as A
from .array import DeviceArray

F = _abi.f

enabled = os.environ.get("SP_STEADY", "1") != "0"
fork_leaf_grads = os.environ.get("SP_STEADY_FORK", "1") != "0"
depth = 0            # > 0 while a root is being evaluated (nested Lazy.run calls are plain)
update_count = 0     # replayed optimiser steps (they write into recycled parameter buffers)
config_epoch = 0     # bumped by anything that changes which kernels an op launches
WARM_STEPS = int(os.environ.get("SP_STEADY_WARM", "3"))
MAX_RECORDS = 8      # recordings per root before it stays eager for good
stats = {"records": 0, "replays_f": 0, "replays_b": 0, "replays_u": 0, "aborted": 0,
         "speculative_b": 0}
# Launch the backward graph right behind the forward graph when the loop has been asking for
# gradients every step (the host then reads the loss through an event on a copy stream).
# Measured on the README loop it buys 3 us of 191 (the loop's critical path is the host batch's
# H2D + forward + sync, not the host gap it removes): off by default.
speculate_backward = os.environ.get("SP_STEADY_SPECULATE", "0") == "1"


def set_enabled(flag: bool) -> None:
    global enabled
    enabled = bool(flag)


def invalidate() -> None:
    """Settings changed (precision, fusion, kernel policy, data-parallel set-up): recordings
    made before hold other kernels."""
    global config_epoch
    config_epoch += 1


class _Record:
    """One recording of a root: the graphs that read ONE set of parameter buffers."""

    __slots__ = ("root_state", "key", "pool", "placeholders", "static_inputs", "static_values",
                 "param_arrays", "shadow_refs", "graph_f", "kernels_f", "root", "epoch",
                 "graph_b", "kernels_b", "grads", "b_epoch", "side", "graph_u", "kernels_u",
                 "u_vars", "u_opt", "u_hyper", "u_outs", "u_refreshed", "user_inputs", "keep",
                 "exchange_bucket", "ready_event", "b_requested", "spec_streak", "pending_b", "mirror", "pending_f")

    def __init__(self):
        self.graph_b = self.graph_u = None
        self.epoch = 0
        self.b_epoch = -1
        self.keep = []
        self.ready_event = None
        self.mirror = None
        self.b_requested = False
        self.spec_streak = 0
        self.pending_b = ()
        self.pending_f = ()


class _RootState:
    __slots__ = ("placeholders", "params", "others", "eligible", "sig", "streak", "records",
                 "n_records", "dead", "pool")

    def __init__(self, root):
        from .core import Variable
        from .graph import Placeholder

        self.placeholders, self.params, self.others = [], [], []
        self.eligible = True
        seen = set()

        def walk(node):
            if id(node) in seen:
                return
            seen.add(id(node))
            if isinstance(node, Placeholder):
                self.placeholders.append(node)
                return
            for child in getattr(node, "arguments", ()):
                if hasattr(child, "run"):
                    walk(child)
                elif isinstance(child, Variable):
                    if id(child) not in seen:
                        seen.add(id(child))
                        self.params.append(child)
                elif isinstance(child, (DeviceArray, np.ndarray)):
                    self.eligible = False  # a raw array baked into the graph: stay eager
                # plain Python values (ints, strings) are part of the function: ignored

        walk(root)
        if not self.placeholders:
            self.eligible = False
        self.sig = None
        self.streak = 0
        self.records = {}
        self.n_records = 0
        self.dead = False
        self.pool = None


def _input_signature(state):
    from .core import Variable

    sig = []
    for ph in state.placeholders:
        args = ph.arguments
        if len(args) != 1:
            return None
        v = args[0]
        if isinstance(v, Variable):
            a = v._array
            if a._thunk is not None or v.local_gradients:
                return None  # the result of an op: its tape would have to be recorded too
            sig.append((0, a.shape, a.dtype.str))
        elif type(v) is DeviceArray:
            if v._thunk is not None:
                return None
            sig.append((1, v.shape, v.dtype.str))
        elif isinstance(v, np.ndarray) and v.dtype.kind in "iu":
            sig.append((2, v.shape, v.dtype.str))  # integer labels / indices
        else:
            return None
    return tuple(sig)


def _eager(root):
    global depth
    depth += 1
    try:
        return root._run()
    finally:
        depth -= 1


def run_root(root):
    """`root.run()` called from user code (not from inside another run)."""
    state = root.__dict__.get("_steady_state")
    if state is None:
        state = root.__dict__["_steady_state"] = _RootState(root)
    if state.dead or not state.eligible or not A.device_ready() or A.capturing():
        return _eager(root)
    from . import distributed

    if distributed.is_initialized() and distributed.world_size() > 1 and not _dp_ok:
        return _eager(root)
    sig = _input_signature(state)
    if sig is None:
        state.streak = 0
        return _eager(root)
    sig = (sig, config_epoch)
    if sig != state.sig:
        if state.sig is not None and state.sig[1] != config_epoch:
            state.records.clear()  # other kernels now
        state.sig = sig
        state.streak = 0
    ptrs = tuple([p._array._ptr for p in state.params])
    if None in ptrs:
        return _eager(root)
    key = (sig, ptrs)
    rec = state.records.get(key)
    if rec is not None:
        if _shadows_intact(rec):
            return _replay_forward(rec)
        del state.records[key]
        state.streak = 0
    state.streak += 1
    if state.streak <= WARM_STEPS or A.capturing():
        return _eager(root)
    if state.n_records >= MAX_RECORDS:
        state.dead = True
        return _eager(root)
    try:
        rec = _record_forward(root, state, key)
    except _Abort:
        stats["aborted"] += 1
        state.dead = True
        return _eager(root)
    return _replay_forward(rec)


_dp_ok = os.environ.get("SP_STEADY_DP", "1") == "1"  # replay with a data-parallel exchange
# Data parallel: issue the gradient exchange inside the backward graph (two parts, on a stream
# of their own beside the sweep) instead of inside the update graph.  Measured on 2 GPUs
# (CIFAR CNN, scripts/steady_parts.py): backward 72 -> 94 us for an update that gets 14 us
# shorter -- the collective is latency-bound and its CTAs queue behind compute kernels that
# fill the chip, so two of them cost more than they hide.  Off by default.
exchange_in_backward = os.environ.get("SP_STEADY_DP_OVERLAP", "0") == "1"


class _Abort(Exception):
    pass


_MIRROR_BYTES = 64
_MIRROR_SLOTS = 16


def _mirror_ring():
    """(page-locked [slots, 17] uint32 array: 16 payload words + the sequence word, per-slot
    generation counters), or None."""
    buf = A.pinned_empty((_MIRROR_SLOTS, 17), np.uint32)
    if buf is None:
        return None
    buf[...] = 0
    return buf, [0] * _MIRROR_SLOTS


def _shadows_intact(rec) -> bool:
    """The TF32 shadows (rounded / split copies) of the parameters that the recorded kernels
    read must still be the arrays' current shadows (U refreshes exactly those)."""
    for arr, rounded, splits in rec.shadow_refs:
        sh = arr._shadow
        if sh is None:
            if rounded is not None or splits:
                return False
            continue
        if sh[0] != A.inplace_epoch or (rounded is not None and sh[1] is not rounded):
            return False
        cur = sh[2]
        for k, v in splits:
            if cur.get(k) is not v:
                return False
    return True


def _snapshot_shadows(arrays):
    refs = []
    for arr in arrays:
        sh = arr._shadow
        if sh is None or sh[0] != A.inplace_epoch:
            refs.append((arr, None, ()))
        else:
            refs.append((arr, sh[1], tuple(sh[2].items())))
    return refs


# --------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------


def _torch():
    import torch

    return torch


def _capture(pool, fn):
    """Record fn() into a CUDA graph allocated from `pool`.  Returns (graph, result, kernels)."""
    torch = _torch()
    graph = torch.cuda.CUDAGraph()
    launches0 = _abi.counter.launches
    A.forbid_uploads = True
    # No garbage collection while the stream is capturing: a cycle collected in the middle of a
    # recording may hold CUDA graphs, streams or pinned buffers of an EARLIER model (a previous
    # root's recordings), and destroying those calls into the driver -- which invalidates the
    # capture in progress ("operation not permitted when stream is capturing").  torch no longer
    # collects before a capture by default, so do it here, then hold the collector off.
    gc_was_on = gc.isenabled()
    gc.collect()
    gc.disable()
    try:
        with A.pool_bypass():
            with torch.cuda.graph(graph, pool=pool):
                result = fn()
    except Exception as exc:  # noqa: BLE001 - anything that cannot be captured: stay eager
        if gc_was_on:
            gc.enable()
        A.forbid_uploads = False
        try:
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
        if os.environ.get("SP_STEADY_DEBUG") == "1":
            raise
        raise _Abort(str(exc)) from exc
    if gc_was_on:
        gc.enable()
    A.forbid_uploads = False
    return graph, result, _abi.counter.launches - launches0


def _record_forward(root, state, key):
    from .capture import _static_like
    from .core import Variable

    torch = _torch()
    if state.pool is None:
        state.pool = torch.cuda.graph_pool_handle()
    rec = _Record()
    rec.root_state = state
    rec.key = key
    rec.pool = state.pool
    rec.placeholders = list(state.placeholders)
    rec.param_arrays = [p._array for p in state.params]
    user_values = [ph.arguments[0] for ph in rec.placeholders]
    rec.static_inputs, rec.static_values = [], []
    with A.pool_bypass():
        for v in user_values:
            src = v._array if isinstance(v, Variable) else v
            static = _static_like(src)
            rec.static_inputs.append(static)
            rec.static_values.append(Variable(static) if isinstance(v, Variable) else static)
    global depth
    for ph, sv in zip(rec.placeholders, rec.static_values):
        ph.arguments = [sv]
    depth += 1
    try:
        def forward():
            r = root._run()
            if isinstance(r, Variable) and r._array._ptr is None and r._array._thunk is not None:
                r._array.ptr  # a deferred root value: launched here, as reading it would
            return r

        rec.graph_f, rec.root, rec.kernels_f = _capture(rec.pool, forward)
    finally:
        depth -= 1
        for ph, uv in zip(rec.placeholders, user_values):
            ph.arguments = [uv]
    if not isinstance(rec.root, Variable) or rec.root._array._thunk is not None:
        raise _Abort("root value is not a launched Variable")
    # Arrays of the recorded tape whose kernel was NOT launched (a consumer fused them away: the
    # first conv's full-resolution output when conv + leaky_relu + pool ran as one kernel) are
    # launched on demand by whoever reads them -- eagerly, after a replay (the deferred input
    # gradient needs that tensor).  Re-armed at every replay: a buffer materialised for one
    # step's inputs must not be mistaken for the next step's.
    from . import core

    pending = []
    for node in core._tape_order(rec.root):
        arr = node._array
        if type(arr) is DeviceArray and arr._ptr is None and arr._thunk is not None:
            pending.append((arr, arr._thunk, arr._tag))
    rec.pending_f = pending
    rec.shadow_refs = _snapshot_shadows(rec.param_arrays)
    state.records[key] = rec
    state.n_records += 1
    stats["records"] += 1
    return rec


def _feed(dst: DeviceArray, value) -> None:
    """Bring a placeholder value into its static buffer."""
    from .capture import _copy_into
    from .core import Variable

    src = value._array if isinstance(value, Variable) else value
    if type(src) is DeviceArray and src._ptr is None and src._thunk is None and src._host is not None:
        # host data that has not been uploaded yet: upload straight into the static buffer,
        # and let the user's array share it (valid until the next run of this root)
        host = src._host
        if host.size:
            if host.dtype != dst.dtype:
                A.upload_f64_narrowed(host, dst.ptr)
            else:
                F.sp_memcpy_h2d(dst.ptr, host.ctypes.data, host.nbytes, A.stream())
                A._counters["h2d_bytes"] += host.nbytes
                if host.ctypes.data in A._pinned.owned:
                    A._pinned.release(host)
                elif host.nbytes >= A._PIN_THRESHOLD:
                    A._keep_alive_until_copied(host)
        src._buf, src._ptr, src._host = dst._buf, dst._ptr, None
        return
    _copy_into(dst, src)


def _replay_forward(rec):
    from .core import Variable

    user_values = [ph.arguments[0] for ph in rec.placeholders]
    resident = []
    for dst, v in zip(rec.static_inputs, user_values):
        src = v._array if isinstance(v, Variable) else v
        if (type(src) is DeviceArray and src._ptr is not None and src.dtype.itemsize == 4
                and src.dtype == dst.dtype and src.shape == dst.shape):
            resident.append((dst, src))  # already on the device: copied below, all in one launch
        else:
            _feed(dst, v)
    if resident:
        n = len(resident)
        F.sp_copy_multi(n, (C.c_void_p * n)(*[s_._ptr for _, s_ in resident]),
                        (C.c_void_p * n)(*[d._ptr for d, _ in resident]),
                        _abi.dims([d.size for d, _ in resident]), A.stream())
    for arr, thunk, tag_ in rec.pending_f:
        if arr._thunk is None:  # materialised after an earlier replay: pending again
            arr._buf = arr._ptr = None
            arr._thunk, arr._tag = thunk, tag_
            arr._tf32 = arr._shadow = None
    rec.graph_f.replay()
    _abi.counter.launches += rec.kernels_f
    stats["replays_f"] += 1
    rec.epoch += 1
    rec.user_inputs = user_values
    static_root = rec.root
    value = static_root._array
    out = DeviceArray.empty(value.shape, value.dtype, value.size)
    out._tf32 = value._tf32
    # the root value is final here: a host read (`np.isnan(loss_val.array)`, README.md:319) waits
    # for this event on a copy stream, not for whatever is enqueued behind it
    if rec.ready_event is None:
        rec.ready_event = A.new_event()
    ring = None
    if 0 < value.nbytes <= _MIRROR_BYTES and value.nbytes % 4 == 0:
        ring = rec.mirror
        if ring is None:
            ring = rec.mirror = _mirror_ring() or False
    if ring:
        # a small result (the loss): the kernel that makes the caller's private copy also
        # publishes it to a page-locked host slot (payload, then a sequence word behind a system
        # fence), so a host read polls one word instead of synchronising a stream -- no copy,
        # no event wait, no wake-up.  A ring of slots: results may be read a few steps later.
        buf, gens = ring
        slot = rec.epoch % _MIRROR_SLOTS
        gen = gens[slot] = (gens[slot] + 1) & 0x7fffffff or 1
        view = buf[slot]
        F.sp_copy_publish(value._ptr, out._ptr, value.nbytes, view.ctypes.data, gen, A.stream())
        F.sp_event_record(rec.ready_event, A.stream())
        out._ready = (rec.ready_event, view, gens, slot, gen)
    else:
        if value.size:
            F.sp_memcpy_d2d(out._ptr, value._ptr, value.nbytes, A.stream())
        F.sp_event_record(rec.ready_event, A.stream())
        out._ready = rec.ready_event
    if rec.graph_b is not None and speculate_backward:
        # The loop has asked for the gradients of every recent result of this recording: start
        # the backward pass now, behind the forward pass, instead of after the host has read
        # the loss and come back (the GPU would idle for the sync wake-up plus ~20 us of
        # Python).  B only writes its own recorded buffers and is idempotent, so a result whose
        # gradients are never requested costs device time, nothing else; one such miss stops
        # the speculation for this recording until the pattern is back.
        rec.spec_streak = rec.spec_streak + 1 if rec.b_requested else 0
        rec.b_requested = False
        if rec.spec_streak >= 2:
            rec.graph_b.replay()
            _abi.counter.launches += rec.kernels_b
            stats["replays_b"] += 1
            stats["speculative_b"] += 1
            rec.b_epoch = rec.epoch
    fresh = Variable.__new__(Variable)
    fresh._array = out
    fresh.local_gradients = static_root.local_gradients
    fresh._serial = static_root._serial
    fresh._steady = (rec, rec.epoch, update_count)
    return fresh


# --------------------------------------------------------------------------------------------
# backward
# --------------------------------------------------------------------------------------------


def gradients_of(variable, tag):
    """`get_gradients(variable)` for a Variable returned by a replayed run, or None to take the
    eager path."""
    from . import core

    rec, epoch, updates = tag
    if rec.epoch != epoch or update_count - updates > 1:
        # (the second update after a forward pass writes into the parameter buffers it read)
        raise RuntimeError(
            "get_gradients: this result belongs to an earlier run() of its lazy graph; the "
            "steady-state replay has since overwritten the tape's buffers (call get_gradients "
            "before the next run(), or sp.set_steady_replay(False))")
    state = rec.root_state
    if state.sig is None or state.sig[1] != config_epoch or state.records.get(rec.key) is not rec:
        return None
    if rec.graph_b is None:
        try:
            _record_backward(rec)
        except _Abort:
            stats["aborted"] += 1
            state.dead = True
            state.records.pop(rec.key, None)
            return None
    rec.b_requested = True
    for arr, thunk, tag_ in rec.pending_b:
        if arr._thunk is None:  # materialised after an earlier replay: pending again
            arr._buf = arr._ptr = None
            arr._thunk, arr._tag = thunk, tag_
            arr._tf32 = arr._shadow = None
    if rec.b_epoch != epoch:  # (else: already launched speculatively behind the forward pass)
        rec.graph_b.replay()
        _abi.counter.launches += rec.kernels_b
        stats["replays_b"] += 1
        rec.b_epoch = epoch
    # this step's view of the recorded gradients: the same buffers, with the recorded root and
    # input Variables replaced by the ones the caller holds (one C-level dict copy, then the
    # few keys that differ)
    static = rec.grads
    out = core.Gradients()
    dict.update(out, dict.items(static))  # (the raw items: Gradients' own views materialise)
    swaps = [(rec.root, variable)]
    for sv, uv in zip(rec.static_values, rec.user_inputs):
        if isinstance(sv, core.Variable) and isinstance(uv, core.Variable):
            swaps.append((sv, uv))
    for sv, uv in swaps:
        if dict.__contains__(out, sv):
            dict.__setitem__(out, uv, dict.pop(out, sv))
    if static._deferred:
        deferred = {k: list(v) for k, v in static._deferred.items()}
        for sv, uv in swaps:
            if sv in deferred:
                deferred[uv] = deferred.pop(sv)
        out._deferred = deferred
    out._steady = (rec, epoch, updates)
    out._reduced = static._reduced
    return out


def _count_leaf_jobs(root):
    """Tape edges into learnable leaves reachable from `root` (= closures `fork` will run)."""
    from . import core

    n = 0
    for node in core._tape_order(root):
        for child, _ in node.local_gradients:
            if not child.local_gradients and getattr(child, "is_learnable", False):
                n += 1
    return n


def _record_backward(rec):
    from . import core
    from . import distributed as dp

    torch = _torch()
    side = rec.side = torch.cuda.Stream() if fork_leaf_grads else None
    forked = []  # (variable, gradient array) in production order
    # Data parallel: the exchange moves from the optimiser step into this graph, on the side
    # stream, in two parts -- everything but the gradient produced last (the first layer's) is
    # packed and all-reduced while the main stream still computes the rest of the sweep, the
    # last one right after its kernel -- so only a small latency-bound collective is exposed.
    comm = None
    if (side is not None and dp.is_initialized() and dp.world_size() > 1
            and dp._state["backend"] == "nccl" and exchange_in_backward):
        comm = dp._direct_nccl()
    n_jobs = _count_leaf_jobs(rec.root) if comm is not None else 0
    exchanged = {}  # variable -> view of the bucket holding the summed gradient
    bucket = [None, 0]  # [DeviceArray, floats used]

    comm_stream = torch.cuda.Stream() if comm is not None else None

    def exchange(pairs):
        """Pack the gradients of `pairs` (complete on the side stream by now) into the next
        bucket slots and sum them over the ranks, on a stream of their own: the side stream
        goes on with the remaining parameter gradients."""
        import ctypes as C

        pairs = [(v, g) for v, g in pairs if v not in exchanged]
        if not pairs:
            return
        sizes = [g.size for _, g in pairs]
        layout = dp.bucket_layout(sizes)
        base_off = bucket[1]
        base = bucket[0]._ptr + 4 * base_off
        tab = (C.c_void_p * len(pairs))(*[g.ptr for _, g in pairs])
        ready = torch.cuda.Event()
        ready.record(side)
        comm_stream.wait_event(ready)
        with torch.cuda.stream(comm_stream):
            F.sp_pack_multi(len(pairs), tab, _abi.dims(sizes), _abi.dims(layout[:-1]), base,
                            A.stream())
            F.sp_allreduce_sum_f32(comm, base, layout[-1], A.stream())
        for (v, g), off in zip(pairs, layout):
            exchanged[v] = DeviceArray(g.shape, A.FLOAT, bucket[0]._buf, base + 4 * off)
        bucket[1] = base_off + layout[-1]

    def fork(fn, path_value):
        """Run a parameter-gradient closure on the side stream (a branch of the graph)."""
        if (type(path_value) is DeviceArray and path_value._ptr is None
                and not (fn.fuses is not None and fn.fuses(path_value))):
            path_value.ptr  # produced on the main stream: both branches read it
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        F.sp_set_scratch_lane(1)
        try:
            with torch.cuda.stream(side):
                if comm is not None and len(forked) == n_jobs - 1 and n_jobs > 1:
                    # every gradient but the last is complete: exchange them now (a variable
                    # that also receives the last contribution waits for the second part)
                    exchange([(v, g) for v, g in _latest(forked) if v is not fn.child])
                value = fn(path_value)
                if type(value) is DeviceArray and value._ptr is None:
                    value.ptr
        finally:
            F.sp_set_scratch_lane(0)
        forked.append((fn.child, value))
        return value

    def body():
        core._leaf_fork = fork if side is not None else None
        try:
            grads = core._get_gradients(rec.root)
        finally:
            core._leaf_fork = None
        if forked:
            if comm is not None:
                exchange(_latest(forked))
                for v, view in exchanged.items():
                    dict.__setitem__(grads, v, view)
                grads._reduced = True
                done_comm = torch.cuda.Event()
                done_comm.record(comm_stream)
                torch.cuda.current_stream().wait_event(done_comm)
            done = torch.cuda.Event()
            done.record(side)
            torch.cuda.current_stream().wait_event(done)
        return grads

    if comm is not None:
        total = sum((p._array.size + 3) // 4 * 4 for p in rec.root_state.params
                    if getattr(p, "is_learnable", False))
        with A.pool_bypass():
            bucket[0] = DeviceArray.empty((total + 64,), A.FLOAT, total + 64)
        rec.keep.append(bucket[0])
    rec.graph_b, rec.grads, rec.kernels_b = _capture(rec.pool, body)
    rec.keep = []  # (the update step fills it with the gradient arrays it expects)
    # Arrays of the recorded sweep whose kernel has NOT been launched (a consumer fused them
    # away, e.g. the pooled gradient the first layer's filter gradient expands itself) are
    # launched on demand by whoever reads them -- eagerly, after a replay.  Re-arm them at
    # every replay: a buffer materialised for one step must not be mistaken for the next's.
    pending = []
    seen = set()
    candidates = list(dict.values(rec.grads))
    for entries in rec.grads._deferred.values():
        candidates.extend(pv for _, pv in entries)
    for arr in candidates:
        if (type(arr) is DeviceArray and id(arr) not in seen and arr._ptr is None
                and arr._thunk is not None):
            seen.add(id(arr))
            pending.append((arr, arr._thunk, arr._tag))
    rec.pending_b = pending
    if comm is not None:
        rec.exchange_bucket = bucket[0]


def _latest(forked):
    """[(variable, its gradient array after the last contribution so far)], production order."""
    last = {}
    for v, g in forked:
        last[v] = g
    return list(last.items())


# --------------------------------------------------------------------------------------------
# update
# --------------------------------------------------------------------------------------------


def optimiser_step(opt, variables, gradients, tag) -> bool:
    """`opt.training_step(variables, gradients)` with a gradient dict from a replayed backward
    pass.  True: done (replayed); False: the caller runs the eager update."""
    global update_count
    rec, epoch, updates = tag
    if rec.epoch != epoch or rec.b_epoch != epoch or update_count != updates:
        return False
    state = rec.root_state
    if state.records.get(rec.key) is not rec or state.sig[1] != config_epoch:
        return False
    if rec.graph_u is None:
        if A.capturing():
            return False
        if not _u_eligible(rec, variables):
            state.dead = True  # this loop updates something the recordings are not keyed by
            state.records.clear()
            return False
        try:
            _record_update(rec, opt, variables, gradients)
        except _Abort:
            stats["aborted"] += 1
            state.dead = True
            state.records.pop(rec.key, None)
            return False
    elif (opt is not rec.u_opt or opt._hyper() != rec.u_hyper or len(variables) != len(rec.u_vars)
          or any(a is not b for a, b in zip(variables, rec.u_vars))):
        return False
    for v, g in zip(rec.u_vars, rec.keep):
        if dict.get(gradients, v) is not g:
            return False  # somebody replaced a gradient in the dict: eager update
    rec.graph_u.replay()
    update_count += 1
    _abi.counter.launches += rec.kernels_u
    stats["replays_u"] += 1
    rec.b_epoch = -1  # one update per backward pass
    for v, out, refreshed in zip(rec.u_vars, rec.u_outs, rec.u_refreshed):
        v._array = out
        sh = out._shadow
        if sh is not None and sh[0] == A.inplace_epoch:
            # shadows the graph does not refresh describe the array's previous contents
            has_round, kinds = refreshed
            if sh[1] is not None and not has_round:
                sh[1] = None
            if len(sh[2]) != len(kinds):
                for k in [k for k in sh[2] if k not in kinds]:
                    del sh[2][k]
        out._aux = None
    return True


def _u_eligible(rec, variables) -> bool:
    """Every updated variable must be one of the parameters the recording is keyed by (its
    buffer identity is what selects the recording next time)."""
    params = rec.root_state.params
    ids = {id(p) for p in params}
    if len({id(v) for v in variables}) != len(variables):
        return False
    if any(id(v) not in ids for v in variables):
        return False
    return all(v._array._ptr is not None for v in variables)


def _record_update(rec, opt, variables, gradients):
    state = rec.root_state
    params = state.params
    index = {id(p): i for i, p in enumerate(params)}
    # write into the parameter arrays of another recording of this root (the buffers of two
    # steps ago: nothing reads them any more), else into new static arrays
    donor = None
    for other in state.records.values():
        if other is rec or other.key[0] != rec.key[0]:
            continue  # (recordings of other input shapes read the same parameters)
        taken = any(r.graph_u is not None and r is not other and
                    all(o is other.param_arrays[index[id(v)]] for o, v in zip(r.u_outs, r.u_vars))
                    for r in state.records.values())
        if not taken and all(other.param_arrays[index[id(v)]].shape == v._array.shape
                             for v in variables):
            donor = other
            break
    if donor is not None:
        outs = [donor.param_arrays[index[id(v)]] for v in variables]
    else:
        with A.pool_bypass():
            outs = [DeviceArray.empty(v._array.shape, A.FLOAT, v._array.size) for v in variables]
    olds = [v._array for v in variables]
    # the shadows the update (re)computes for the arrays it writes: those its inputs have
    refreshed = []
    for old in olds:
        want_round, kinds = A.tf32_shadow_requests(old)
        refreshed.append((bool(want_round), frozenset(kinds)))
    # eager once: creates the optimiser state (moments, step counters) outside the capture --
    # no: that would advance the step.  State creation is host-side only (zeros are launched),
    # so create it here explicitly.
    opt._ensure_state(variables)
    _torch().cuda.synchronize()

    def body():
        opt._update(variables, gradients, outs)

    try:
        rec.graph_u, _, rec.kernels_u = _capture(rec.pool, body)
    finally:
        for v, old in zip(variables, olds):
            v._array = old  # the capture only recorded: the replay below applies the step
    rec.u_vars = list(variables)
    rec.u_opt = opt
    rec.u_hyper = opt._hyper()
    rec.u_outs = outs
    rec.keep = [dict.get(gradients, v) for v in variables]
    rec.u_refreshed = refreshed
