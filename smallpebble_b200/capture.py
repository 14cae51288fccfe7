"""CUDA-graph capture of a steady-state step.

The README training loop is launch- and Python-bound on a B200 (the CIFAR CNN step is ~45
kernel launches of 5-30 us each).  `capture(fn)` records ONE call of an ordinary SmallPebble
step function -- the user's own code: `Placeholder.assign_value`, `loss.run()`,
`sp.get_gradients`, `adam.training_step` -- into a CUDA graph and replays it for every
later batch:

    def train_step(x, y):                       # x, y arrive as device arrays
        X_in.assign_value(sp.Variable(x)); y_true.assign_value(y)
        loss_val = loss.run()
        adam.training_step(learnables, sp.get_gradients(loss_val))
        return loss_val

    step = sp.capture(train_step)
    for xbatch, ybatch in sp.batch(X, y, 128):
        loss_val = step(xbatch, ybatch)         # H2D into static buffers + one graph launch

Nothing about the ops changes: the same C-ABI kernels are launched, on the capture stream,
into torch's graph-private memory pool.  What makes the step capturable is that no launch
parameter depends on the step number (Adam's bias correction reads a device-resident
counter) and that the fp32 / int32 input buffers are static.  Shapes must not change between
calls (a new shape triggers a new capture).
"""

from __future__ import annotations

import numpy as np

from . import _abi
from . import array as A
from .array import DeviceArray

F = _abi.f


def _static_like(value) -> DeviceArray:
    """A persistent device buffer able to hold `value` (floats -> fp32, ints -> int32)."""
    if isinstance(value, DeviceArray):
        return DeviceArray.empty(value.shape, value.dtype)
    a = np.asarray(value)
    dtype = np.int32 if a.dtype.kind in "iub" else np.float32
    return DeviceArray.empty(a.shape, dtype)


def _copy_into(dst: DeviceArray, value) -> None:
    if isinstance(value, DeviceArray):
        if value.shape != dst.shape or value.dtype != dst.dtype:
            raise ValueError("captured step: input shape/dtype changed")
        F.sp_memcpy_d2d(dst.ptr, value.ptr, dst.nbytes, A.stream())
        return
    staged = DeviceArray.from_host(value, dtype=dst.dtype)  # pinned staging + conversion
    if staged.shape != dst.shape:
        raise ValueError(f"captured step: input shape {staged.shape} != captured {dst.shape}")
    host = staged._host
    if not host.size:
        return
    st = A.stream()
    if host.dtype != dst.dtype:
        # caller's pinned float64 batch: raw DMA on the copy stream, narrowed into the static
        # buffer on the compute stream (in front of the graph launch)
        A.upload_f64_narrowed(host, dst.ptr)
        return
    F.sp_memcpy_h2d(dst.ptr, host.ctypes.data, host.nbytes, st)
    A._counters["h2d_bytes"] += host.nbytes
    if host.ctypes.data in A._pinned.owned:
        A._pinned.release(host)
    else:
        A._keep_alive_until_copied(host)


class CapturedStep:
    """Callable returned by `capture(fn)`."""

    def __init__(self, fn, warmup: int = 2):
        self.fn = fn
        self.warmup = max(int(warmup), 1)
        self._graphs = {}  # input signature -> (graph, static inputs, outputs)

    @staticmethod
    def _signature(args):
        sig = []
        for a in args:
            if isinstance(a, DeviceArray):
                sig.append((a.shape, str(a.dtype)))
            else:
                a = np.asarray(a)
                sig.append((a.shape, "i" if a.dtype.kind in "iub" else "f"))
        return tuple(sig)

    def _record(self, args):
        import torch

        A.require_cuda()
        with A.pool_bypass():  # static buffers, warm-up and capture allocate from torch itself
            return self._record_unpooled(args, torch)

    def _record_unpooled(self, args, torch):
        static = [_static_like(a) for a in args]
        for dst, a in zip(static, args):
            _copy_into(dst, a)
        # eager warm-up on a side stream: lazy allocations (Adam state, library scratch,
        # kernel attributes, NCCL channels) must exist before capture starts
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self.fn(*static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        launches0 = _abi.counter.launches
        # (no garbage collection inside the capture: destroying an older graph or stream that a
        # collected cycle holds would invalidate it -- see steady._capture)
        import gc

        gc_was_on = gc.isenabled()
        gc.collect()
        gc.disable()
        try:
            with torch.cuda.graph(graph):
                outputs = self.fn(*static)
        finally:
            if gc_was_on:
                gc.enable()
        kernels = _abi.counter.launches - launches0
        return graph, static, outputs, kernels

    def __call__(self, *args):
        sig = self._signature(args)
        entry = self._graphs.get(sig)
        if entry is None:
            entry = self._graphs[sig] = self._record(args)
            # the recording pass consumed `args` as its warm-up/capture batch and did not run
            # the captured work on it: fall through and replay so this call's update is applied
        graph, static, outputs, kernels = entry
        for dst, a in zip(static, args):
            _copy_into(dst, a)
        graph.replay()
        _abi.counter.launches += kernels  # kernels of OUR library executed by the replay
        return outputs

    @property
    def kernels_per_replay(self) -> int:
        return max((e[3] for e in self._graphs.values()), default=0)


def capture(fn, warmup: int = 2) -> CapturedStep:
    """Wrap a step function so that its device work is replayed from a CUDA graph.

    The first call with a given set of input shapes runs `fn` `warmup` times eagerly and once
    under capture (so optimiser state advances `warmup` steps on that first batch -- call it
    before the timed / real training loop if that matters), then replays."""
    return CapturedStep(fn, warmup)
