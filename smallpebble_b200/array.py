"""DeviceArray: the thing `Variable.array` holds instead of a NumPy ndarray.

A thin, always-contiguous fp32 (or int/uint8 for indices) device buffer.  PyTorch is used
for exactly two things here: its caching allocator (`torch.empty`) and the current CUDA
stream handle.  Every arithmetic method launches a hand-written kernel through the C ABI
(`_abi.f.*`); nothing is computed by torch and nothing falls back to the CPU.

Host data handed to `Variable(np_array)` is uploaded lazily, on first device use, so that
graph construction (`sp.linearlayer`, `sp.Placeholder`, `sp.get_learnables`, ...) works on a
machine without a GPU; touching `.ptr` without CUDA raises.
"""

from __future__ import annotations

import ctypes as C
import functools
import math
import numbers
import os
import weakref

import numpy as np

from . import _abi

F = _abi.f
FLOAT = np.dtype(np.float32)
_DEVICE_DTYPES = {np.dtype(np.float32), np.dtype(np.int32), np.dtype(np.int64), np.dtype(np.uint8)}

_torch = None
_device = None


def _lazy_torch():
    global _torch
    if _torch is None:
        import torch

        _torch = torch
    return _torch


def require_cuda():
    """The product path has no CPU fallback: fail loudly when there is no GPU/extension."""
    global _device
    if _device is not None:
        return _device
    torch = _lazy_torch()
    if not torch.cuda.is_available():
        raise _abi.SmallPebbleCudaError(
            "smallpebble_b200 needs a CUDA device (B200, sm_100a): torch.cuda.is_available() is "
            "False. There is no CPU fallback."
        )
    _abi.load()
    torch.cuda.init()
    _device = torch.device("cuda", torch.cuda.current_device())
    # bind our statically linked CUDA runtime to torch's context on this device
    torch.empty(1, device=_device)
    global _raw_stream, _device_index
    _device_index = _device.index
    _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if _raw_stream is None:  # older torch: slower, same handle
        _raw_stream = lambda idx: torch.cuda.current_stream(idx).cuda_stream  # noqa: E731
    for hook in _ready_hooks:
        hook()
    return _device


_ready_hooks = []


def on_device_ready(hook) -> None:
    """Run `hook()` once the CUDA library is bound to a device (library-global settings)."""
    _ready_hooks.append(hook)
    if _device is not None:
        hook()


def device_ready() -> bool:
    return _device is not None


def reset_device_cache():
    """Call after torch.cuda.set_device() (one process per GPU: smallpebble_b200.distributed)."""
    global _device, _raw_stream
    _device = None
    _raw_stream = None
    trim_pool()  # recycled buffers belong to the previous device
    _upload_ring.reset()
    _prestaged.clear()


_raw_stream = None
_device_index = 0


def stream() -> int:
    """Raw cudaStream_t of torch's current stream on our device (all kernels launch here)."""
    if _raw_stream is None:
        require_cuda()
    return _raw_stream(_device_index)


def _host_dtype(a: np.ndarray) -> np.dtype:
    """Device dtype for a host array: all real data is fp32; bools -> uint8; ints stay ints
    only when explicitly requested (labels / indices go through `from_host(..., dtype)`)."""
    if a.dtype == np.bool_:
        return np.dtype(np.uint8)
    return FLOAT


# Free list in front of torch's caching allocator.  A training step allocates the same ~35
# buffer sizes over and over; recycling the torch buffers by exact byte size turns each
# `DeviceArray.empty` from a ~2 us torch.empty call into a stack pop (the README loop is
# host-bound on a B200).  A recyclable buffer is held through a Block object (`_buf`): arrays
# and views share the Block, and when the LAST reference to it dies the Block's destructor
# hands the torch tensor to the per-size stack instead of freeing it.  Everything runs on one
# stream, so reuse is stream-ordered; the pool is bypassed while a CUDA graph is being recorded
# (graph-private memory must come from torch's capture pool).  The Block type and the stacks
# live in C (csrc/fastcall.c: no Python-level __del__ on the hot path); `_PyBlock` below is the
# same thing in Python for when that module is not built.
_POOL_MAX_BLOCK = 64 << 20
_POOL_MAX_TOTAL = 2 << 30
_pool_bypass = 0


class _PyBlock:
    __slots__ = ("tensor", "ptr", "nbytes", "recycle")

    def __init__(self, tensor, ptr, nbytes, recycle):
        self.tensor, self.ptr, self.nbytes, self.recycle = tensor, ptr, nbytes, recycle

    def __del__(self):
        try:
            if self.recycle and not _pool_bypass and _py_pool["bytes"] + self.nbytes <= _POOL_MAX_TOTAL:
                _py_pool["blocks"].setdefault(self.nbytes, []).append((self.tensor, self.ptr))
                _py_pool["bytes"] += self.nbytes
        except Exception:  # interpreter shutdown: module globals may be gone
            pass


_py_pool = {"blocks": {}, "bytes": 0}


def _py_pool_alloc(nbytes):
    if _pool_bypass:
        return None
    bucket = _py_pool["blocks"].get(nbytes)
    if not bucket:
        return None
    tensor, ptr = bucket.pop()
    _py_pool["bytes"] -= nbytes
    return _PyBlock(tensor, ptr, nbytes, 1)


def _py_pool_trim():
    _py_pool["blocks"].clear()
    _py_pool["bytes"] = 0


def _py_pool_stats():
    return sum(len(v) for v in _py_pool["blocks"].values()), _py_pool["bytes"]


_fast = _abi._load_fast()
if _fast is not None and hasattr(_fast, "pool_alloc"):
    _pool_alloc, _block_new = _fast.pool_alloc, _fast.block_new
    _pool_trim, _pool_stats, _pool_set_bypass = _fast.pool_trim, _fast.pool_stats, _fast.pool_bypass
else:
    _pool_alloc, _block_new = _py_pool_alloc, _PyBlock
    _pool_trim, _pool_stats = _py_pool_trim, _py_pool_stats
    _pool_set_bypass = lambda delta: None  # noqa: E731 - the Python pool reads _pool_bypass


def pool_stats():
    """(buffers, bytes) currently waiting in the free list."""
    return _pool_stats()


class pool_bypass:
    """Context manager: allocate straight from torch (used around CUDA-graph capture)."""

    def __enter__(self):
        global _pool_bypass
        _pool_bypass += 1
        _pool_set_bypass(1)

    def __exit__(self, *exc):
        global _pool_bypass
        _pool_bypass -= 1
        _pool_set_bypass(-1)


def capturing() -> bool:
    """True while a CUDA graph is being recorded (sp.capture): buffers must be static."""
    return _pool_bypass > 0


# Deferred arrays (their pending kernel may read a PARAMETER buffer: conv2d / matmul outputs
# waiting for a fusing consumer).  An optimiser that updates parameters in place launches them
# first (graph._settle_before_inplace_update); the out-of-place Adam never needs to.
_pending_refs = []
_weakref = weakref.ref


def flush_pending_parameter_readers() -> None:
    """Launch every deferred array that is still pending (in creation order)."""
    refs = list(_pending_refs)
    del _pending_refs[:]
    for ref in refs:
        arr = ref()
        if arr is not None and arr._ptr is None and arr._thunk is not None:
            arr._materialize()


def trim_pool() -> None:
    """Hand every recycled buffer back to torch's allocator."""
    _pool_trim()
    _scalar_ones.clear()


_read_side = {}


def _read_stream() -> int:
    """A non-blocking stream of our own for early host reads (one per device)."""
    st = _read_side.get(_device_index)
    if st is None:
        slot = (C.c_void_p * 1)()
        F.sp_stream_create(slot)
        st = _read_side[_device_index] = slot[0]
    return st


def new_event() -> int:
    slot = (C.c_void_p * 1)()
    F.sp_event_create(slot)
    return slot[0]


forbid_uploads = False  # set by steady.py around its graph captures


_scalar_ones = {}  # shape -> read-only device constant 1.0 (the gradient seed of a scalar loss)


def scalar_one(shape):
    """(array of ones, fresh?) for a one-element `shape`.  Outside CUDA-graph recording the
    constant is created once and shared (saves a fill launch per get_gradients); callers must
    treat a shared one as read-only."""
    arr = _scalar_ones.get(shape)
    if arr is None:
        if _pool_bypass:  # recording: a constant created now would live in the graph's pool
            arr = DeviceArray.full(shape, 1.0)
            arr._aux = ONES
            return arr, True
        arr = _scalar_ones[shape] = DeviceArray.full(shape, 1.0)
        arr._aux = ONES
    return arr, False


ONES = ("ones",)  # `_aux` marker of an all-ones gradient seed (closures may special-case it)


_new_array = object.__new__
_np_dtype = type(FLOAT)
_prod = math.prod


class DeviceArray:
    __slots__ = ("_buf", "_ptr", "shape", "dtype", "size", "_host", "_thunk", "_tag", "_aux",
                 "_tf32", "_shadow", "_ready", "__weakref__")
    __array_priority__ = 1000.0

    def __init__(self, shape, dtype=FLOAT, buf=None, ptr=None, host=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = math.prod(self.shape)
        self._buf = buf
        self._ptr = ptr
        self._host = host
        self._thunk = None
        self._tag = None
        self._aux = None  # by-product of the producing kernel, e.g. ("colsum0", column sums)
        self._tf32 = None    # True: every value is TF32-representable (rounded by its producer)
        self._shadow = None  # [epoch, rounded copy, {(kind, block): split copy}] (tf32_shadows)
        self._ready = None   # event after which the contents are final (steady.py: read early)

    # ---- construction ---------------------------------------------------------------------
    @staticmethod
    def empty(shape, dtype=FLOAT, size=None) -> "DeviceArray":
        """Uninitialised device array.  `size` (= prod(shape)) may be passed by callers that
        already know it -- this is the most frequent call of a training step."""
        dev = _device if _device is not None else require_cuda()
        self = _new_array(DeviceArray)
        if type(shape) is not tuple:
            shape = tuple(shape)
        self.shape = shape
        if dtype is not FLOAT and type(dtype) is not _np_dtype:
            dtype = np.dtype(dtype)
        self.dtype = dtype
        if size is None:
            size = _prod(shape)
        self.size = size
        self._host = None
        self._thunk = None
        self._tag = None
        self._aux = None
        self._tf32 = None
        self._shadow = None
        self._ready = None
        nbytes = (size or 1) * dtype.itemsize
        if _pool_bypass or nbytes > _POOL_MAX_BLOCK:
            self._buf = buf = _torch.empty(nbytes, dtype=_torch.uint8, device=dev)
            self._ptr = buf.data_ptr()
        else:
            blk = _pool_alloc(nbytes)
            if blk is None:
                buf = _torch.empty(nbytes, dtype=_torch.uint8, device=dev)
                blk = _block_new(buf, buf.data_ptr(), nbytes, 1)
            self._buf = blk  # shared by views; the last holder's death recycles the buffer
            self._ptr = blk.ptr
        return self

    @staticmethod
    def empty_like(other) -> "DeviceArray":
        "float32 array of `other`'s shape (gradient buffers)."
        return DeviceArray.empty(other.shape, FLOAT, other.size)

    @staticmethod
    def deferred(shape, thunk, tag, size=None) -> "DeviceArray":
        """A float32 array whose producing kernel has not been launched yet.  `thunk(out)` fills
        a freshly allocated `out`; it runs the first time anything needs the data (`.ptr`,
        `.numpy()`, ...).  `tag` describes the pending op so that a consumer able to fuse it
        (core.maxpool2d after core.leaky_relu, and their gradients) can skip it altogether --
        the fusion across the API seam of SURVEY.md 8(f)."""
        self = _new_array(DeviceArray)
        self.shape = shape if type(shape) is tuple else tuple(shape)
        self.dtype = FLOAT
        self.size = _prod(self.shape) if size is None else size
        self._buf = None
        self._ptr = None
        self._host = None
        self._aux = None
        self._tf32 = None
        self._shadow = None
        self._ready = None
        self._thunk = thunk
        self._tag = tag
        if len(_pending_refs) > 512:  # drop the dead and the launched
            _pending_refs[:] = [r for r in _pending_refs
                                if (a := r()) is not None and a._ptr is None]
        _pending_refs.append(_weakref(self))
        return self

    @property
    def pending_tag(self):
        """The tag of a not-yet-launched deferred array, else None."""
        return self._tag if self._ptr is None else None

    def _materialize(self):
        thunk = self._thunk
        out = DeviceArray.empty(self.shape, FLOAT, self.size)
        thunk(out)
        self._adopt(out)

    def _adopt(self, out):
        """A deferred array takes over `out`'s buffer: whoever calls this fills it instead of
        the pending thunk (a consumer that fused the producing op into its own kernel)."""
        self._buf, self._ptr = out._buf, out._ptr
        self._tf32 = out._tf32
        self._thunk = None
        self._tag = None

    @staticmethod
    def from_host(a, dtype=None) -> "DeviceArray":
        """Wrap host data; the copy to the device happens on first use.  The host values are
        snapshotted now (converted to the device dtype, into pinned memory when a GPU is
        present so the later host->device copy is a true async DMA)."""
        a = np.asarray(a)
        dtype = np.dtype(dtype) if dtype is not None else _host_dtype(a)
        if dtype not in _DEVICE_DTYPES:
            raise TypeError(f"unsupported device dtype {dtype}")
        nbytes = a.size * dtype.itemsize
        if nbytes >= _PIN_THRESHOLD and _pinned.available():
            if (dtype == FLOAT and a.dtype in _DIRECT_DTYPES and a.flags.c_contiguous
                    and _is_pinned(a)):
                # The caller's batch already sits in page-locked memory: no host pass at all.
                # The upload DMAs the raw bytes from it and float64 is narrowed on the device.
                # Like the reference's Variable(array), this aliases the caller's array until
                # the copy has run (the array is kept alive; do not overwrite it in between).
                if eager_staging and a.dtype == np.float64 and a.size and _device is not None:
                    _prestage(a)  # the DMA starts now, on the copy stream
                return DeviceArray(a.shape, dtype, host=a)
            host = _pinned.take(a.shape, dtype)
            np.copyto(host, a, casting="unsafe")  # one pass: convert + stage
            out = DeviceArray(host.shape, dtype, host=host)
            weakref.finalize(out, _pinned.abandon, host.ctypes.data)  # never uploaded: recycle
            return out
        # snapshot: the caller may mutate its array after Variable(...).  (np.array, not
        # np.ascontiguousarray: the latter promotes 0-d arrays to shape (1,).)
        host = np.array(a, dtype=dtype, order="C", copy=True)
        return DeviceArray(host.shape, dtype, host=host)

    @staticmethod
    def full(shape, value, dtype=FLOAT) -> "DeviceArray":
        out = DeviceArray.empty(shape, dtype)
        if out.dtype != FLOAT:
            raise TypeError("full() supports float32 only")
        F.sp_fill_f32(out._ptr, out.size, float(value), stream())
        return out

    @staticmethod
    def zeros(shape, dtype=FLOAT) -> "DeviceArray":
        return DeviceArray.full(shape, 0.0, dtype)

    # ---- device pointer (uploads pending host data) ---------------------------------------
    @property
    def ptr(self) -> int:
        if self._ptr is None:
            self._upload()
        return self._ptr

    def _upload(self):
        if self._thunk is not None:
            self._materialize()
            return
        if forbid_uploads:
            # steady.py is recording this call into a CUDA graph: host data created inside the
            # step (a random mask, say) would be baked into the graph.  Refuse: the recording is
            # abandoned and the step keeps running eagerly.
            raise RuntimeError("host data uploaded inside a step that is being recorded")
        dev = require_cuda()
        host = self._host
        if host.dtype != self.dtype:
            # pinned float64 source: raw DMA (copy stream), then narrow on the device
            buf = _torch.empty(host.size * 4, dtype=_torch.uint8, device=dev)
            upload_f64_narrowed(host, buf.data_ptr())
            self._buf, self._ptr, self._host = buf, buf.data_ptr(), None
            return
        nbytes = max(host.size, 1) * host.dtype.itemsize
        buf = _torch.empty(nbytes, dtype=_torch.uint8, device=dev)
        if host.size:
            # pinned staging buffer: async DMA, the buffer returns to the pool once the copy
            # has completed; pageable memory: the runtime stages it before returning
            F.sp_memcpy_h2d(buf.data_ptr(), host.ctypes.data, host.nbytes, stream())
            _counters["h2d_bytes"] += host.nbytes
            if host.ctypes.data in _pinned.owned:
                _pinned.release(host)
            elif host.nbytes >= _PIN_THRESHOLD:
                _keep_alive_until_copied(host)  # direct DMA from the caller's pinned array
        self._buf, self._ptr, self._host = buf, buf.data_ptr(), None

    @property
    def on_device(self) -> bool:
        return self._ptr is not None

    # ---- NumPy interop -------------------------------------------------------------------
    def numpy(self) -> np.ndarray:
        """Device -> host copy (synchronises the stream)."""
        if self._ptr is None and self._thunk is not None:
            self._materialize()
        if self._ptr is None:
            return np.array(self._host, dtype=self.dtype, copy=True)
        out = np.empty(self.shape, self.dtype)
        ready = self._ready
        if type(ready) is tuple:
            # (ready event, host slot, slot generations, slot, generation): the producer also
            # published the contents to a page-locked host slot -- payload words, then the
            # generation in word 16 behind a system fence (sp_copy_publish).  Poll that word:
            # no copy, no stream or event wait -- unless the ring has come round to this slot
            # since, or the word does not show up (then the event path below)
            ready, mirror, gens, slot, gen = ready
            if gens[slot] == gen:
                spins = 0
                while mirror[16] != gen and spins < 2000000:
                    spins += 1
                if mirror[16] == gen:
                    out.reshape(-1).view(np.uint32)[:] = mirror[:out.nbytes // 4]
                    _counters["d2h_bytes"] += out.nbytes
                    return out
        if ready is not None:
            # the contents were final at this event (the result of a replayed forward pass):
            # read them on the copy stream as soon as it has fired, without waiting for work
            # enqueued behind it on the compute stream (steady.py's speculative backward pass)
            side = _read_stream()
            F.sp_stream_wait_event(side, ready)
            F.sp_memcpy_d2h(out.ctypes.data, self._ptr, out.nbytes, side)
        else:
            F.sp_memcpy_d2h(out.ctypes.data, self._ptr, out.nbytes, stream())
        _counters["d2h_bytes"] += out.nbytes
        return out

    get = numpy

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method == "__call__" and not kwargs and ufunc in _UFUNC_TO_OP and len(inputs) == 2:
            return _binary_any(_UFUNC_TO_OP[ufunc], inputs[0], inputs[1])
        host = [np.asarray(x) if isinstance(x, DeviceArray) else x for x in inputs]
        if "out" in kwargs:
            return NotImplemented
        return getattr(ufunc, method)(*host, **kwargs)

    def item(self):
        return self.numpy().item()

    def __float__(self):
        return float(self.numpy().reshape(-1)[0]) if self.size == 1 else float(self.numpy())

    def __int__(self):
        return int(float(self))

    def __bool__(self):
        if self.size != 1:
            raise ValueError("The truth value of an array with more than one element is ambiguous.")
        return bool(float(self))

    def __len__(self):
        if not self.shape:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def __repr__(self):
        where = "device" if self._ptr is not None else "host-pending"
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, {where})"

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    # ---- views / copies --------------------------------------------------------------------
    def reshape(self, *shape) -> "DeviceArray":
        if len(shape) == 1 and not isinstance(shape[0], numbers.Integral):
            shape = tuple(shape[0])
        shape = _reshape_plan(shape, self.size)
        if self._ptr is None and self._thunk is not None:
            self._materialize()
        if self._ptr is None:
            return DeviceArray(shape, self.dtype, host=self._host.reshape(shape))
        out = _new_array(DeviceArray)  # a view: shares the buffer (and its Block)
        out.shape = shape
        out.dtype = self.dtype
        out.size = self.size
        out._buf = self._buf
        out._ptr = self._ptr
        out._host = None
        out._thunk = None
        out._tag = None
        out._aux = None
        out._tf32 = self._tf32
        out._shadow = None
        out._ready = self._ready
        return out

    def copy(self) -> "DeviceArray":
        out = DeviceArray.empty(self.shape, self.dtype)
        F.sp_memcpy_d2d(out._ptr, self.ptr, self.nbytes, stream())
        return out

    def astype(self, dtype, copy=True):
        if np.dtype(dtype) == self.dtype:
            return self.copy() if copy else self
        return DeviceArray.from_host(self.numpy().astype(dtype), dtype=dtype)

    def swapaxes(self, a, b) -> "DeviceArray":
        return swapaxes(self, a, b)

    @property
    def T(self):
        return transpose(self, tuple(reversed(range(self.ndim))))

    def sum(self, axis=None, keepdims=False):
        return reduce_sum(self, axis, keepdims)

    def __getitem__(self, index):
        return take(self, index)

    # ---- arithmetic (all on device) ------------------------------------------------------
    def __add__(self, o):
        return _binary_any("add", self, o)

    def __radd__(self, o):
        return _binary_any("add", o, self)

    def __sub__(self, o):
        return _binary_any("sub", self, o)

    def __rsub__(self, o):
        return _binary_any("sub", o, self)

    def __mul__(self, o):
        return _binary_any("mul", self, o)

    def __rmul__(self, o):
        return _binary_any("mul", o, self)

    def __truediv__(self, o):
        return _binary_any("div", self, o)

    def __rtruediv__(self, o):
        return _binary_any("div", o, self)

    def __neg__(self):
        return unary("neg", self)

    def __pow__(self, e):
        if e == 2:
            return unary("square", self)
        if e == 1:
            return self.copy()
        if e == 0.5:
            return unary("sqrt", self)
        raise NotImplementedError("DeviceArray ** e supports e in {0.5, 1, 2}")

    def _rebind(self, other: "DeviceArray"):
        self._buf, self._ptr, self._host = other._buf, other.ptr, None
        self._tf32, self._shadow, self._ready = other._tf32, None, None
        return self

    def __iadd__(self, o):
        return self._rebind(_binary_any("add", self, o))

    def __isub__(self, o):
        return self._rebind(_binary_any("sub", self, o))

    def __imul__(self, o):
        return self._rebind(_binary_any("mul", self, o))

    def __itruediv__(self, o):
        return self._rebind(_binary_any("div", self, o))


class _PinnedPool:
    """Recycled page-locked staging buffers for host->device uploads (torch.empty(pin_memory)
    as the allocator).  A buffer handed out by take() is owned by one pending DeviceArray;
    release() records a CUDA event after the H2D copy and the buffer is reused only once
    that event has completed."""

    def __init__(self):
        self.free = {}      # nbytes -> [tensor]
        self.owned = {}     # id(view base) -> (nbytes, tensor)
        self.in_flight = []  # (event, nbytes, tensor)
        self._ok = None

    def available(self) -> bool:
        if self._ok is None:
            torch = _lazy_torch()
            self._ok = bool(torch.cuda.is_available())
        return self._ok

    def _reclaim(self):
        keep = []
        for ev, nbytes, t in self.in_flight:
            if ev.query():
                self.free.setdefault(nbytes, []).append(t)
            else:
                keep.append((ev, nbytes, t))
        self.in_flight = keep

    def take(self, shape, dtype) -> np.ndarray:
        torch = _lazy_torch()
        nbytes = math.prod(shape) * dtype.itemsize
        self._reclaim()
        bucket = self.free.get(nbytes)
        t = bucket.pop() if bucket else torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        view = t.numpy().view(dtype).reshape(shape)
        self.owned[view.ctypes.data] = (nbytes, t)
        return view

    def abandon(self, address: int):
        entry = self.owned.pop(address, None)
        if entry is not None:
            self.free.setdefault(entry[0], []).append(entry[1])

    def release(self, host: np.ndarray):
        entry = self.owned.pop(host.ctypes.data, None)
        if entry is None:
            return
        ev = _lazy_torch().cuda.Event()
        ev.record()
        self.in_flight.append((ev, entry[0], entry[1]))


_pinned = _PinnedPool()
_PIN_THRESHOLD = 64 * 1024
_DIRECT_DTYPES = {np.dtype(np.float32), np.dtype(np.float64)}
_in_flight_sources = []  # (event, ndarray): caller arrays an async H2D is still reading


class _UploadRing:
    """Raw host->device uploads of page-locked float64 batches on a dedicated COPY stream, so
    that the DMA of step i+1's batch overlaps the kernels of step i that are still queued on
    the compute stream (the README loop calls `sp.Variable(xbatch)` while the previous step's
    backward pass is in flight).  Three library-owned staging buffers, reused round-robin:
    the copy stream waits for the event recorded after a slot's last consumer (the narrowing
    kernel on the compute stream) before overwriting it, and the compute stream waits for the
    copy's event before reading it.  The buffers never go back to torch's allocator while in
    use, so no cross-stream reuse hazard exists.  Stream and events are raw handles driven
    through the C ABI (an upload is on the README loop's critical path: a torch.cuda.Event
    costs more host time than the copy call), two events per slot, re-recorded."""

    SLOTS = 3

    def __init__(self):
        self.stream = None
        self.slots = [{"buf": None, "ptr": 0, "nbytes": 0, "landed": None, "free": None,
                       "free_valid": False, "host": None, "gen": 0, "stage": None, "f32": False}
                      for _ in range(self.SLOTS)]
        self.next = 0

    def reset(self):
        self.__init__()

    def upload(self, host: np.ndarray, wait: bool = True):
        """Start the DMA of `host` into the next slot; `wait`: order the compute stream behind
        it now (else the caller does, later: `slot["landed"]`).  The slot keeps `host` alive
        until it is reused, which first makes sure that copy has completed."""
        if self.stream is None:
            st = (C.c_void_p * 1)()
            F.sp_stream_create(st)
            self.stream = st[0]
        slot = self.slots[self.next]
        self.next = (self.next + 1) % self.SLOTS
        slot["gen"] += 1  # (a pre-staged upload of an older generation is gone)
        nbytes = host.nbytes
        if slot["landed"] is None:
            slot["landed"], slot["free"] = new_event(), new_event()
        elif slot["host"] is not None:
            if slot["f32"]:
                F.sp_h2d_f64_as_f32_finish()        # (its copies are all enqueued)
            F.sp_event_synchronize(slot["landed"])  # three uploads ago: long complete
        if slot["nbytes"] < nbytes:
            torch = _lazy_torch()
            if slot["free_valid"]:
                F.sp_event_synchronize(slot["free"])  # rare: the slot grows; its last reader must be done
            slot["buf"] = torch.empty(nbytes, dtype=torch.uint8, device=require_cuda())
            slot["ptr"] = slot["buf"].data_ptr()
            slot["nbytes"] = nbytes
            slot["free_valid"] = False
            # the new block may be a recycled one: order the copy after everything queued so far
            F.sp_event_record(slot["free"], stream())
            F.sp_stream_wait_event(self.stream, slot["free"])
        if slot["free_valid"]:
            F.sp_stream_wait_event(self.stream, slot["free"])
        if host_pipe and host.dtype == np.float64 and nbytes >= _HOST_PIPE_MIN:
            # narrowed on host cores while the copy runs: half the bytes cross PCIe, the slot
            # ends up holding fp32 (csrc/host_pipe.cu); returns at once
            stage = slot["stage"]
            if stage is None or stage.size < host.size:
                stage = slot["stage"] = pinned_empty((host.size,), np.float32)
            if stage is not None:
                F.sp_h2d_f64_as_f32_begin(host.ctypes.data, stage.ctypes.data, slot["ptr"],
                                          host.size, slot["landed"], self.stream)
                slot["f32"] = True
                slot["host"] = host
                if wait:
                    self.ready(slot)
                return slot
        F.sp_memcpy_h2d(slot["ptr"], host.ctypes.data, nbytes, self.stream)
        F.sp_event_record(slot["landed"], self.stream)
        slot["f32"] = False
        slot["host"] = host
        if wait:
            F.sp_stream_wait_event(stream(), slot["landed"])
        return slot

    @staticmethod
    def ready(slot):
        """Order the compute stream behind the slot's copy."""
        if slot["f32"]:
            F.sp_h2d_f64_as_f32_finish()  # the pool has enqueued every chunk and the event
        F.sp_stream_wait_event(stream(), slot["landed"])

    @staticmethod
    def consumed(slot):
        F.sp_event_record(slot["free"], stream())  # after the kernel that read the slot
        slot["free_valid"] = True


_upload_ring = _UploadRing()


# Uploads whose DMA was started when the caller wrapped its page-locked float64 batch
# (`sp.Variable(xbatch)`), before the array was first needed on the device: host data pointer ->
# (slot, slot generation, bytes).  In the README loop the batch is wrapped a few API calls before
# `run()` reaches the upload, and the host-to-device copy is on the loop's critical path.
_prestaged = {}
eager_staging = os.environ.get("SP_EAGER_UPLOAD", "1") != "0"
# Optional: float64 batches of a megabyte or more narrowed to fp32 by a pool of host threads
# while the copy is in flight (csrc/host_pipe.cu), which halves the bytes on the longest leg of
# the README loop's critical path.  OFF by default: on this pool's hosts (16 virtual cores)
# eight threads need ~100 us for the 3.1 MB batch -- the copy engine moves the float64 bytes in
# 61 us -- so the loop got slower, not faster (scripts/e2e_phases.py).
host_pipe = os.environ.get("SP_HOST_PIPE", "0") == "1"
_HOST_PIPE_MIN = 1 << 20


def _prestage(host: np.ndarray) -> None:
    if forbid_uploads or capturing():
        return
    if len(_prestaged) >= 8:  # wrapped but never used: their slots have been recycled
        for k in [k for k, (sl, gen, _) in _prestaged.items() if sl["gen"] != gen]:
            del _prestaged[k]
        if len(_prestaged) >= 8:
            return
    slot = _upload_ring.upload(host, wait=False)
    _prestaged[host.ctypes.data] = (slot, slot["gen"], host.nbytes)


def upload_f64_narrowed(host: np.ndarray, dst_ptr: int) -> None:
    """float64 page-locked `host` -> fp32 at `dst_ptr`: raw DMA on the copy stream (already
    under way when the array was pre-staged), narrowing kernel on the compute stream."""
    pre = _prestaged.pop(host.ctypes.data, None) if _prestaged else None
    if pre is not None and pre[0]["gen"] == pre[1] and pre[2] == host.nbytes:
        slot = pre[0]
        _UploadRing.ready(slot)
    else:
        slot = _upload_ring.upload(host)
    if slot["f32"]:  # narrowed on the way (host pipeline): fp32 arrived
        F.sp_memcpy_d2d(dst_ptr, slot["ptr"], host.size * 4, stream())
        _counters["h2d_bytes"] += host.size * 4
    else:
        F.sp_cast_f64_f32(slot["ptr"], dst_ptr, host.size, stream())
        _counters["h2d_bytes"] += host.nbytes
    _UploadRing.consumed(slot)


def pinned_empty(shape, dtype):
    """A fresh NumPy array in page-locked memory (None when there is no CUDA device).  Used by
    `sp.batch`: a minibatch gathered into it is uploaded by DMA straight from the array the
    user holds (see `from_host`), and it is an ordinary ndarray in every other respect -- the
    pinned block goes back to torch's host allocator when the array is garbage collected."""
    if not _pinned.available():
        return None
    torch = _lazy_torch()
    dtype = np.dtype(dtype)
    nbytes = max(math.prod(shape), 1) * dtype.itemsize
    try:
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    except RuntimeError:
        return None
    return t.numpy()[:math.prod(shape) * dtype.itemsize].view(dtype).reshape(shape)


def _is_pinned(a: np.ndarray) -> bool:
    flag = C.c_int(0)
    F.sp_host_is_pinned(a.ctypes.data, flag)
    return bool(flag.value)


def _keep_alive_until_copied(host: np.ndarray) -> None:
    """Hold a reference to a PAGE-LOCKED source of an asynchronous host-to-device copy until the
    copy has run: a ring of eight raw events, the oldest waited for (long complete) when it comes
    round.  Pageable sources need none of this: cudaMemcpyAsync stages them before it returns."""
    if host.nbytes < _PIN_THRESHOLD and host.ctypes.data not in _pinned.owned:
        return  # from_host pins nothing below the threshold
    if len(_in_flight_sources) < 8:
        ev = new_event()
    else:
        ev, _ = _in_flight_sources.pop(0)
        F.sp_event_synchronize(ev)
    F.sp_event_record(ev, stream())
    _in_flight_sources.append((ev, host))

_UFUNC_TO_OP = {np.add: "add", np.subtract: "sub", np.multiply: "mul", np.true_divide: "div"}
_counters = {"h2d_bytes": 0, "d2h_bytes": 0}


def transfer_counters():
    return _counters


def asarray(x) -> DeviceArray:
    if isinstance(x, DeviceArray):
        return x
    return DeviceArray.from_host(x)


@functools.lru_cache(maxsize=4096)
def _reshape_plan(shape, size):
    """The validated target shape of a reshape (the same few every training step)."""
    return _resolve_shape(shape, size)


def _resolve_shape(shape, size):
    shape = [int(s) for s in shape]
    if shape.count(-1) > 1:
        raise ValueError("can only specify one unknown dimension")
    if -1 in shape:
        known = -math.prod(shape)
        if known == 0 or size % known:
            raise ValueError(f"cannot reshape array of size {size} into shape {tuple(shape)}")
        shape[shape.index(-1)] = size // known
    if math.prod(shape) != size:
        raise ValueError(f"cannot reshape array of size {size} into shape {tuple(shape)}")
    return tuple(shape)


def _contig_strides(shape):
    out, s = [], 1
    for n in reversed(shape):
        out.append(s)
        s *= n
    return tuple(reversed(out))


@functools.lru_cache(maxsize=4096)
def _broadcast_plan(shape_a, shape_b):
    """(out_shape, ndim, c_shape, c_stride_a, c_stride_b) with stride 0 on broadcast axes."""
    nd = max(len(shape_a), len(shape_b))
    if nd > _abi.MAX_DIMS:
        raise ValueError(f"arrays with more than {_abi.MAX_DIMS} dimensions are not supported")
    ea = (1,) * (nd - len(shape_a)) + shape_a
    eb = (1,) * (nd - len(shape_b)) + shape_b
    out = []
    for p, q in zip(ea, eb):
        if p != q and p != 1 and q != 1:
            raise ValueError(f"could not broadcast shapes {shape_a} {shape_b}")
        out.append(max(p, q) if 0 not in (p, q) else 0)
    sa = tuple(0 if n == 1 and o != 1 else s for n, o, s in zip(ea, out, _contig_strides(ea)))
    sb = tuple(0 if n == 1 and o != 1 else s for n, o, s in zip(eb, out, _contig_strides(eb)))
    return tuple(out), nd, _abi.dims(out), _abi.dims(sa), _abi.dims(sb)


def unary(op: str, a: DeviceArray, param: float = 0.0) -> DeviceArray:
    out = DeviceArray.empty(a.shape)
    F.sp_ewise_unary(_abi.UNARY[op], a.ptr, out._ptr, a.size, param, stream())
    return out


def binary(op: str, a: DeviceArray, b: DeviceArray) -> DeviceArray:
    """NumPy-broadcasting elementwise op between two device arrays."""
    shape, nd, c_shape, c_sa, c_sb = _broadcast_plan(a.shape, b.shape)
    out = DeviceArray.empty(shape)
    F.sp_ewise_binary(_abi.BINARY[op], a.ptr, c_sa, b.ptr, c_sb, out._ptr, nd, c_shape, stream())
    return out


def _binary_any(op, a, b):
    """Binary op where one side may be a Python scalar or host array."""
    a_dev, b_dev = isinstance(a, DeviceArray), isinstance(b, DeviceArray)
    if a_dev and isinstance(b, numbers.Real):
        b = float(b)
        if op == "add":
            return unary("add_scalar", a, b)
        if op == "sub":
            return unary("add_scalar", a, -b)
        if op == "mul":
            return unary("scale", a, b)
        if op == "div":
            return unary("scale", a, 1.0 / b) if b != 0 else binary(op, a, asarray(np.float32(b)))
    if b_dev and isinstance(a, numbers.Real):
        a = float(a)
        if op == "add":
            return unary("add_scalar", b, a)
        if op == "mul":
            return unary("scale", b, a)
        if op == "div":
            return unary("recip", b, a)
        if op == "sub":
            return unary("add_scalar", unary("neg", b), a)
    if not a_dev:
        a = asarray(np.asarray(a))
    if not b_dev:
        b = asarray(np.asarray(b))
    return binary(op, a, b)


def accumulate_(dst: DeviceArray, src: DeviceArray) -> DeviceArray:
    """dst += src in place (same shape)."""
    F.sp_accumulate(dst.ptr, src.ptr, dst.size, stream())
    dst._aux = None  # by-products described the old contents
    dst._tf32 = None  # a sum of TF32 values is not a TF32 value
    dst._shadow = None
    dst._ready = None
    return dst


def _norm_axes(axis, ndim):
    if axis is None:
        return tuple(range(ndim))
    if isinstance(axis, numbers.Integral):
        axis = (axis,)
    axes = sorted({int(ax) + ndim if ax < 0 else int(ax) for ax in axis})
    for ax in axes:
        if not 0 <= ax < ndim:
            raise ValueError(f"axis {ax} is out of bounds for array of dimension {ndim}")
    return tuple(axes)


@functools.lru_cache(maxsize=1024)
def _reduce_plan(shape, axis, keepdims):
    """((outer, red, inner) per kernel, final shape, any axis reduced?) of np.sum(a, axis)."""
    axes = _norm_axes(axis, len(shape))
    shape = list(shape)
    # fold runs from the last axis backwards so earlier axis numbers stay valid
    runs, i = [], len(axes) - 1
    while i >= 0:
        hi = lo = axes[i]
        while i > 0 and axes[i - 1] == lo - 1:
            i -= 1
            lo = axes[i]
        runs.append((lo, hi))
        i -= 1
    steps = []
    for lo, hi in runs:
        steps.append((math.prod(shape[:lo]), math.prod(shape[lo:hi + 1]), math.prod(shape[hi + 1:])))
        shape[lo:hi + 1] = [1] * (hi + 1 - lo)
    if keepdims:
        final = tuple(shape)
    else:
        final = tuple(n for i, n in enumerate(shape) if i not in axes)
    return tuple(steps), final, bool(axes)


def reduce_sum(a: DeviceArray, axis=None, keepdims=False) -> DeviceArray:
    """np.sum(a, axis): runs of adjacent reduced axes are folded one kernel at a time."""
    if axis is not None and not isinstance(axis, (tuple, numbers.Integral)):
        axis = tuple(int(ax) for ax in axis)  # lists / arrays of axes: hashable for the cache
    steps, final, any_axis = _reduce_plan(a.shape, axis, bool(keepdims))
    if not any_axis:
        return a.copy()
    cur = a
    for outer, red, inner in steps:
        out = DeviceArray.empty((outer, inner))
        if outer * inner:
            F.sp_reduce_sum(cur.ptr, out._ptr, outer, red, inner, stream())
        cur = out
    return cur.reshape(final)


def max_all(a: DeviceArray) -> DeviceArray:
    out = DeviceArray.empty(())
    F.sp_reduce_max_all(a.ptr, out._ptr, a.size, stream())
    return out


def strided_copy(src: DeviceArray, src_strides, src_offset, dst: DeviceArray, dst_strides,
                 dst_offset, shape):
    """dst[dst_offset + i.dst_strides] = src[src_offset + i.src_strides] over `shape`."""
    if len(shape) > _abi.MAX_DIMS:
        raise ValueError(f"arrays with more than {_abi.MAX_DIMS} dimensions are not supported")
    if math.prod(shape) == 0:
        return
    F.sp_strided_copy(src.ptr + 4 * src_offset, _abi.dims(src_strides),
                      dst.ptr + 4 * dst_offset, _abi.dims(dst_strides), len(shape),
                      _abi.dims(shape), stream())


def transpose(a: DeviceArray, perm) -> DeviceArray:
    perm = tuple(perm)
    shape = tuple(a.shape[p] for p in perm)
    st = _contig_strides(a.shape)
    out = DeviceArray.empty(shape)
    strided_copy(a, tuple(st[p] for p in perm), 0, out, _contig_strides(shape), 0, shape)
    return out


def swapaxes(a: DeviceArray, i, j) -> DeviceArray:
    perm = list(range(a.ndim))
    perm[i], perm[j] = perm[j], perm[i]
    return transpose(a, perm)


def pad(a: DeviceArray, pad_width) -> DeviceArray:
    """np.pad(a, pad_width) with zeros."""
    pw = _norm_pad(pad_width, a.ndim)
    shape = tuple(n + lo + hi for n, (lo, hi) in zip(a.shape, pw))
    out = DeviceArray.zeros(shape)
    ost = _contig_strides(shape)
    offset = sum(lo * s for (lo, _), s in zip(pw, ost))
    strided_copy(a, _contig_strides(a.shape), 0, out, ost, offset, a.shape)
    return out


def unpad(a: DeviceArray, pad_width) -> DeviceArray:
    """a[lo:n-hi, ...] for each axis (gradient of pad, sp.py:325-327)."""
    pw = _norm_pad(pad_width, a.ndim)
    shape = tuple(n - lo - hi for n, (lo, hi) in zip(a.shape, pw))
    ist = _contig_strides(a.shape)
    offset = sum(lo * s for (lo, _), s in zip(pw, ist))
    out = DeviceArray.empty(shape)
    strided_copy(a, ist, offset, out, _contig_strides(shape), 0, shape)
    return out


def _norm_pad(pad_width, ndim):
    pw = np.asarray(pad_width, dtype=np.int64)
    pw = np.broadcast_to(pw, (ndim, 2))
    return [(int(lo), int(hi)) for lo, hi in pw]


_ARANGE_CACHE_BYTES = 16 << 20
_arange_cache = {}  # shape -> arange, at most _ARANGE_CACHE_BYTES in total


def _arange_index(shape):
    """np.arange(prod(shape)).reshape(shape); small ones are cached (a (60000, 784) source would
    pin 376 MB: those are rebuilt on demand and never kept)."""
    shape = tuple(shape)
    arr = _arange_cache.get(shape)
    if arr is None:
        arr = np.arange(math.prod(shape), dtype=np.int64).reshape(shape)
        if arr.nbytes <= _ARANGE_CACHE_BYTES // 4:
            if sum(a.nbytes for a in _arange_cache.values()) + arr.nbytes > _ARANGE_CACHE_BYTES:
                _arange_cache.clear()
            _arange_cache[shape] = arr
    return arr


def flat_index(shape, index) -> np.ndarray:
    """int64 array of flat source offsets selected by the NumPy index expression `index`
    (advanced indexing: integer / boolean arrays; basic indices never come here, see
    basic_view)."""
    return np.asarray(_arange_index(tuple(shape))[index], order="C")


def basic_view(shape, index):
    """(element offset, strides, shape) of `a[index]` as a strided view of a contiguous array of
    `shape`, when `index` is BASIC (ints, slices, Ellipsis, None) -- NumPy's own rule for "this
    is a view".  None for advanced indices.  The device then moves the data straight from the
    view geometry: no index map is built, uploaded or cached."""
    if not isinstance(index, tuple):
        index = (index,)
    for it in index:
        if not (it is None or it is Ellipsis or isinstance(it, (slice, numbers.Integral))):
            return None
    if builtins_sum(1 for it in index if it is Ellipsis) > 1:
        return None
    n_real = builtins_sum(1 for it in index if it is not None and it is not Ellipsis)
    if n_real > len(shape):
        raise IndexError("too many indices for array")
    if Ellipsis in index:
        k = index.index(Ellipsis)
        index = index[:k] + (slice(None),) * (len(shape) - n_real) + index[k + 1:]
    else:
        index = index + (slice(None),) * (len(shape) - n_real)
    src_strides = _contig_strides(shape)
    offset, out_shape, out_strides, axis = 0, [], [], 0
    for it in index:
        if it is None:
            out_shape.append(1)
            out_strides.append(0)
            continue
        dim, st = shape[axis], src_strides[axis]
        if isinstance(it, numbers.Integral):
            i = int(it)
            if i < -dim or i >= dim:
                raise IndexError(f"index {i} is out of bounds for axis {axis} with size {dim}")
            offset += (i % dim) * st
        else:
            start, stop, step = it.indices(dim)
            n = max(0, (stop - start + (step - (1 if step > 0 else -1))) // step)
            offset += start * st if n else 0
            out_shape.append(n)
            out_strides.append(step * st)
        axis += 1
    return offset, tuple(out_strides), tuple(out_shape)


builtins_sum = sum


def gather_flat(a: DeviceArray, flat: np.ndarray) -> DeviceArray:
    idx = DeviceArray.from_host(flat, dtype=np.int64)
    out = DeviceArray.empty(flat.shape)
    if flat.size:
        F.sp_gather_i64(a.ptr, idx.ptr, out._ptr, flat.size, stream())
    return out


def view_gather(a: DeviceArray, view) -> DeviceArray:
    """Contiguous copy of a strided view (offset, strides, shape) of `a`."""
    offset, strides_, shape = view
    out = DeviceArray.empty(shape)
    strided_copy(a, strides_, offset, out, _contig_strides(shape), 0, shape)
    return out


def view_scatter(dst: DeviceArray, view, src: DeviceArray, accumulate: bool = False):
    """dst.view (=|+=) src, src broadcast to the view's shape.  `accumulate`: np.add.at
    semantics (a view may name an element more than once: sliding windows)."""
    offset, strides_, shape = view
    if math.prod(shape) == 0:
        return
    if len(shape) > _abi.MAX_DIMS:
        raise ValueError(f"arrays with more than {_abi.MAX_DIMS} dimensions are not supported")
    # broadcast src by stride 0 (trailing-aligned, NumPy rules)
    sshape = (1,) * (len(shape) - src.ndim) + tuple(src.shape)
    if len(sshape) != len(shape) or any(x != 1 and x != y for x, y in zip(sshape, shape)):
        raise ValueError(f"could not broadcast input array from shape {src.shape} into shape {shape}")
    cst = _contig_strides(sshape)
    sst = tuple(0 if x == 1 else c for x, c in zip(sshape, cst))
    fn = F.sp_strided_add if accumulate else F.sp_strided_copy
    fn(src.ptr, _abi.dims(sst), dst.ptr + 4 * offset, _abi.dims(strides_), len(shape),
       _abi.dims(shape), stream())


def take(a: DeviceArray, index) -> DeviceArray:
    """a[index] for any NumPy index expression: a basic index is a strided copy straight from
    the view geometry; an advanced one gathers through a host-built index map."""
    view = basic_view(a.shape, index)
    if view is not None:
        return view_gather(a, view)
    return gather_flat(a, flat_index(a.shape, index))


def scatter_add_flat(dst: DeviceArray, flat: np.ndarray, src: DeviceArray):
    """np.add.at(dst.ravel(), flat, src) on device."""
    if flat.size == 0:
        return
    idx = DeviceArray.from_host(flat, dtype=np.int64)
    if src.shape != flat.shape:
        src = binary("add", DeviceArray.zeros(flat.shape), src)  # broadcast to the index shape
    F.sp_scatter_add_i64(dst.ptr, idx.ptr, src.ptr, flat.size, stream())


def scatter_set_flat(dst: DeviceArray, flat: np.ndarray, src: DeviceArray):
    if flat.size == 0:
        return
    idx = DeviceArray.from_host(flat, dtype=np.int64)
    scalar = src.size == 1
    if not scalar and src.shape != flat.shape:
        src = binary("add", DeviceArray.zeros(flat.shape), src)
    F.sp_scatter_set_i64(dst.ptr, idx.ptr, src.ptr, flat.size, 1 if scalar else 0, stream())


def where(cond, a: DeviceArray, b: DeviceArray) -> DeviceArray:
    cond = np.asarray(cond, dtype=np.bool_)
    shape = np.broadcast_shapes(cond.shape, a.shape, b.shape)
    nd = len(shape)
    if nd > _abi.MAX_DIMS:
        raise ValueError(f"arrays with more than {_abi.MAX_DIMS} dimensions are not supported")

    def strides(s):
        e = (1,) * (nd - len(s)) + tuple(s)
        return tuple(0 if n == 1 and o != 1 else st
                     for n, o, st in zip(e, shape, _contig_strides(e)))

    c = DeviceArray.from_host(cond, dtype=np.uint8)
    out = DeviceArray.empty(shape)
    if out.size:
        F.sp_where(c.ptr, _abi.dims(strides(cond.shape)), a.ptr, _abi.dims(strides(a.shape)),
                   b.ptr, _abi.dims(strides(b.shape)), out._ptr, nd, _abi.dims(shape), stream())
    return out


# --------------------------------------------------------------------------------------------
# TF32 error compensation (csrc/tf32.cu): rounded / split shadows of contraction operands
# --------------------------------------------------------------------------------------------

inplace_epoch = 0  # bumped by whoever overwrites parameter buffers IN PLACE (sgd_step, a
                   # captured Adam step): shadows made before that are stale


def note_inplace_update() -> None:
    global inplace_epoch
    inplace_epoch += 1


def _shadow_slot(a: DeviceArray):
    sh = a._shadow
    if sh is None or sh[0] != inplace_epoch:
        sh = a._shadow = [inplace_epoch, None, {}]
    return sh


def tf32_round(a: DeviceArray) -> DeviceArray:
    """`a` rounded to the nearest TF32 values (itself when its producer already rounded it):
    the operand form of the single-pass backward contractions.  Cached on the array."""
    if a._tf32:
        return a
    sh = _shadow_slot(a)
    out = sh[1]
    if out is None:
        out = DeviceArray.empty(a.shape, FLOAT, a.size)
        if a.size:
            F.sp_tf32_round(a.ptr, out._ptr, a.size, stream())
        out._tf32 = True
        sh[1] = out
    return out


def tf32_split(a: DeviceArray, kind: int, block_elems: int) -> DeviceArray:
    """The (r, l[, r]) split of `a` for an error-compensated forward contraction, flat
    [a.size / block_elems, parts, block_elems].  Cached on the array."""
    sh = _shadow_slot(a)
    key = (kind, block_elems)
    out = sh[2].get(key)
    if out is None:
        parts = 2 if kind == _abi.SPLIT_RL else 3
        out = DeviceArray.empty((a.size * parts,), FLOAT, a.size * parts)
        if a.size:
            F.sp_tf32_split(a.ptr, out._ptr, a.size, block_elems, kind, stream())
        sh[2][key] = out
    return out


def tf32_shadow_requests(a: DeviceArray):
    """(wants rounded copy?, [(kind, block_elems), ...]) that were made for `a`: the optimiser
    re-creates the same shadows for the array that replaces it, all tensors in one launch."""
    sh = a._shadow
    if sh is None:
        return False, ()
    return sh[1] is not None, tuple(sh[2])


def tf32_shadow_targets(a: DeviceArray, want_round: bool, kinds, refresh: bool):
    """Buffers for the shadows of `a` that a producer of `a` is about to fill itself (the fused
    Adam kernel): (rounded array or None, (split array, kind, block) or None, kinds left over
    for tf32_shadows_multi).  Existing shadows are only handed out again under `refresh`."""
    sh = _shadow_slot(a)
    rounded = None
    if want_round and not a._tf32:
        if sh[1] is None:
            rounded = DeviceArray.empty(a.shape, FLOAT, a.size)
            rounded._tf32 = True
            sh[1] = rounded
        elif refresh:
            rounded = sh[1]
    split, rest = None, []
    for kind, block in kinds:
        sp = sh[2].get((kind, block))
        if sp is not None and not refresh:
            continue
        if split is not None:
            rest.append((kind, block))
            continue
        if sp is None:
            parts = 2 if kind == _abi.SPLIT_RL else 3
            sp = DeviceArray.empty((a.size * parts,), FLOAT, a.size * parts)
            sh[2][(kind, block)] = sp
        split = (sp, kind, block)
    return rounded, split, rest


def tf32_shadows_multi(jobs, refresh: bool = False) -> None:
    """jobs: [(array, want_rounded, [(kind, block_elems), ...])].  Fills the shadow slots of
    every array with ONE launch per 24 tensors (a job with two split kinds takes two rows).
    `refresh`: shadows that already exist are recomputed into their buffers (the array was
    overwritten: steady.py's optimiser step writes into the arrays of two steps ago)."""
    rows = []
    for a, want_round, kinds in jobs:
        if not a.size:
            continue
        sh = _shadow_slot(a)
        rounded = None
        if want_round and not a._tf32:
            if sh[1] is None:
                rounded = DeviceArray.empty(a.shape, FLOAT, a.size)
                rounded._tf32 = True
                sh[1] = rounded
            elif refresh:
                rounded = sh[1]
        first = True
        for kind, block in kinds:
            sp = sh[2].get((kind, block))
            if sp is not None and not refresh:
                continue
            if sp is None:
                parts = 2 if kind == _abi.SPLIT_RL else 3
                sp = DeviceArray.empty((a.size * parts,), FLOAT, a.size * parts)
                sh[2][(kind, block)] = sp
            rows.append((a, rounded if first else None, sp, block, kind))
            first = False
        if first and rounded is not None:
            rows.append((a, rounded, None, 1, 0))
    if not rows:
        return
    n = len(rows)
    vp = C.c_void_p
    src = (vp * n)(*[r[0].ptr for r in rows])
    rnd = (vp * n)(*[(r[1]._ptr if r[1] is not None else None) for r in rows])
    spl = (vp * n)(*[(r[2]._ptr if r[2] is not None else None) for r in rows])
    numel = (C.c_int64 * n)(*[r[0].size for r in rows])
    blocks = (C.c_int64 * n)(*[r[3] for r in rows])
    kinds = (C.c_int32 * n)(*[r[4] for r in rows])
    F.sp_tf32_shadow_multi(n, src, rnd, spl, numel, blocks, kinds, stream())
