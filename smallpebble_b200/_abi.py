"""ctypes binding of include/smallpebble_b200.h -- the only place Python crosses into the
CUDA library.  Raw device pointers (ints) and a cudaStream_t in, int status out.

There is NO fallback: if the shared library cannot be loaded (or built) every op raises.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libsmallpebble_b200.so")

PRECISION_FP32 = 0
PRECISION_TF32 = 1
PRECISION_TF32X3 = 2      # error-compensated, raw operands (kernels that split in registers)
PRECISION_TF32_SPLIT = 3  # error-compensated, operands pre-split by sp_tf32_split
SPLIT_RL, SPLIT_RLR, SPLIT_RRL = 0, 1, 2
X3_NATIVE, X3_SPLIT2, X3_CONCAT3 = 0, 2, 3

UNARY = {"exp": 0, "log": 1, "neg": 2, "square": 3, "scale": 4, "add_scalar": 5, "recip": 6,
         "sqrt": 7, "copy": 8}
BINARY = {"add": 0, "sub": 1, "mul": 2, "div": 3, "div_grad_b": 4, "exp_mul": 5}

MAX_DIMS = 8

i32, i64, f32, vp, sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t
p_i64 = C.POINTER(C.c_int64)


class ConvGeom(C.Structure):
    """struct sp_conv_geom (include/smallpebble_b200.h)."""

    _fields_ = [(n, C.c_int32) for n in
                ("N", "H", "W", "C", "KH", "KW", "K", "sy", "sx", "pad_t", "pad_l", "OH", "OW")]


# name -> (restype, argtypes); must list EVERY export of the header (tests check this)
SIGNATURES = {
    "sp_abi_version": (C.c_int, []),
    "sp_last_error": (C.c_char_p, []),
    "sp_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "sp_debug_set_umma_policy": (C.c_int, [C.c_int]),
    "sp_debug_tma2_trace": (C.c_int, [vp]),
    "sp_debug_peer_trace": (C.c_int, [vp, vp]),
    "sp_memcpy_h2d": (C.c_int, [vp, vp, sz, vp]),
    "sp_host_is_pinned": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "sp_memcpy_d2h": (C.c_int, [vp, vp, sz, vp]),
    "sp_memcpy_d2d": (C.c_int, [vp, vp, sz, vp]),
    "sp_fill_f32": (C.c_int, [vp, i64, f32, vp]),
    "sp_cast_f64_f32": (C.c_int, [vp, vp, i64, vp]),
    "sp_cast_i64_i32": (C.c_int, [vp, vp, i64, vp]),
    "sp_l2_flush": (C.c_int, [vp, sz, vp]),
    "sp_set_scratch_lane": (C.c_int, [C.c_int]),
    "sp_ewise_unary": (C.c_int, [C.c_int, vp, vp, i64, f32, vp]),
    "sp_ewise_binary": (C.c_int, [C.c_int, vp, p_i64, vp, p_i64, vp, C.c_int, p_i64, vp]),
    "sp_accumulate": (C.c_int, [vp, vp, i64, vp]),
    "sp_leaky_relu_fwd": (C.c_int, [vp, vp, i64, f32, vp]),
    "sp_leaky_relu_bwd": (C.c_int, [vp, vp, vp, i64, f32, vp]),
    "sp_where": (C.c_int, [vp, p_i64, vp, p_i64, vp, p_i64, vp, C.c_int, p_i64, vp]),
    "sp_strided_copy": (C.c_int, [vp, p_i64, vp, p_i64, C.c_int, p_i64, vp]),
    "sp_strided_add": (C.c_int, [vp, p_i64, vp, p_i64, C.c_int, p_i64, vp]),
    "sp_reduce_sum": (C.c_int, [vp, vp, i64, i64, i64, vp]),
    "sp_reduce_max_all": (C.c_int, [vp, vp, i64, vp]),
    "sp_argmax_rows_fwd": (C.c_int, [vp, vp, vp, i64, i64, vp]),
    "sp_argmax_rows_bwd": (C.c_int, [vp, vp, vp, i64, i64, vp]),
    "sp_gather_i64": (C.c_int, [vp, vp, vp, i64, vp]),
    "sp_scatter_add_i64": (C.c_int, [vp, vp, vp, i64, vp]),
    "sp_scatter_set_i64": (C.c_int, [vp, vp, vp, i64, i64, vp]),
    "sp_softmax_fwd": (C.c_int, [vp, vp, i64, i64, i64, vp]),
    "sp_softmax_bwd": (C.c_int, [vp, vp, vp, i64, i64, i64, vp]),
    "sp_xent_fwd": (C.c_int, [vp, vp, vp, i64, i64, vp]),
    "sp_xent_bwd": (C.c_int, [vp, vp, vp, vp, i64, i64, vp]),
    "sp_softmax_xent_fwd": (C.c_int, [vp, vp, vp, vp, i64, i64, vp]),
    "sp_softmax_xent_bwd": (C.c_int, [vp, vp, vp, vp, i64, i64, vp]),
    "sp_softmax_xent_fwd_grad": (C.c_int, [vp, vp, vp, vp, vp, vp, i64, i64, vp]),
    "sp_stream_create": (C.c_int, [vp]),
    "sp_event_create": (C.c_int, [vp]),
    "sp_event_record": (C.c_int, [vp, vp]),
    "sp_event_synchronize": (C.c_int, [vp]),
    "sp_host_pipe_threads": (C.c_int, [C.POINTER(C.c_int)]),
    "sp_h2d_f64_as_f32_begin": (C.c_int, [vp, vp, vp, i64, vp, vp]),
    "sp_h2d_f64_as_f32_finish": (C.c_int, []),
    "sp_copy_publish": (C.c_int, [vp, vp, sz, vp, C.c_uint, vp]),
    "sp_memcpy_d2h_async": (C.c_int, [vp, vp, sz, vp]),
    "sp_stream_wait_event": (C.c_int, [vp, vp]),
    "sp_softmax_xent_bwd_colsum": (C.c_int, [vp, vp, vp, vp, vp, i64, i64, vp]),
    "sp_gemm_bias": (C.c_int, [i64, i64, i64, vp, vp, vp, vp, C.c_int, vp]),
    "sp_gemm_bias_leaky": (C.c_int, [i64, i64, i64, vp, vp, vp, vp, C.c_int, f32, vp]),
    "sp_gemm_bias_softmax_xent": (C.c_int, [i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp]),
    "sp_maxpool2d_fwd": (C.c_int, [vp, vp, vp] + [C.c_int] * 12 + [vp]),
    "sp_maxpool2d_bwd": (C.c_int, [vp, vp, vp] + [C.c_int] * 12 + [vp]),
    "sp_maxpool2d_leaky_fwd": (C.c_int, [vp, vp, vp] + [C.c_int] * 12 + [f32, vp]),
    "sp_maxpool2d_fwd_split": (C.c_int, [vp, vp, vp, vp, i64, C.c_int] + [C.c_int] * 13 + [f32, vp]),
    "sp_maxpool2d_leaky_bwd": (C.c_int, [vp, vp, vp, vp] + [C.c_int] * 12 + [f32, vp]),
    "sp_maxpool2d_leaky_bwd_y": (C.c_int, [vp, vp, vp, vp] + [C.c_int] * 12 + [f32, vp]),
    "sp_gemm": (C.c_int, [C.c_int, C.c_int, i64, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64,
                          i64, i64, C.c_int, vp]),
    "sp_conv2d_fprop": (C.c_int, [vp, vp, vp, C.POINTER(ConvGeom), C.c_int, vp]),
    "sp_conv2d_fprop_leaky": (C.c_int, [vp, vp, vp, C.POINTER(ConvGeom), C.c_int, f32, vp]),
    "sp_conv2d_dgrad": (C.c_int, [vp, vp, vp, C.POINTER(ConvGeom), C.c_int, vp]),
    "sp_conv2d_dgrad_leaky": (C.c_int, [vp, vp, vp, vp, C.POINTER(ConvGeom), C.c_int, f32, vp]),
    "sp_conv2d_dgrad_unpool_supported": (C.c_int, [C.POINTER(ConvGeom), C.c_int, C.c_int, C.c_int,
                                                   C.POINTER(C.c_int)]),
    "sp_conv2d_dgrad_unpool_leaky": (C.c_int, [vp, vp, vp, vp, vp, C.POINTER(ConvGeom), C.c_int,
                                               C.c_int, C.c_int, f32, vp]),
    "sp_conv2d_wgrad_workspace": (sz, [C.POINTER(ConvGeom), C.c_int]),
    "sp_conv2d_wgrad": (C.c_int, [vp, vp, vp, C.POINTER(ConvGeom), C.c_int, vp, sz, vp]),
    "sp_conv2d_wgrad_pooled_supported": (C.c_int, [C.POINTER(ConvGeom), C.c_int, C.POINTER(C.c_int)]),
    "sp_conv2d_leaky_pool2_supported": (C.c_int, [C.POINTER(ConvGeom), C.c_int, C.POINTER(C.c_int)]),
    "sp_conv2d_leaky_pool2_fwd": (C.c_int, [vp, vp, vp, vp, vp, C.POINTER(ConvGeom), C.c_int, f32, vp]),
    "sp_conv2d_wgrad_pooled": (C.c_int, [vp, vp, vp, vp, vp, C.POINTER(ConvGeom), f32, C.c_int, vp,
                                         sz, vp]),
    "sp_adam_multi": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                C.POINTER(vp), p_i64, C.POINTER(vp), f32, f32, f32, f32, vp]),
    "sp_adam_multi_out": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                    C.POINTER(vp), C.POINTER(vp), p_i64, C.POINTER(vp), f32, f32,
                                    f32, f32, vp]),
    "sp_adam_multi_shadow": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                       C.POINTER(vp), C.POINTER(vp), p_i64, C.POINTER(vp),
                                       C.POINTER(vp), C.POINTER(vp), p_i64, C.POINTER(C.c_int32),
                                       f32, f32, f32, f32, vp]),
    "sp_peer_alloc": (C.c_int, [sz, vp, vp]),
    "sp_peer_open": (C.c_int, [vp, vp]),
    "sp_peer_close": (C.c_int, [vp]),
    "sp_peer_free": (C.c_int, [vp]),
    "sp_peer_exchange_create": (C.c_int, [C.c_int, C.c_int, i64, C.POINTER(vp), vp]),
    "sp_peer_exchange_destroy": (C.c_int, [vp]),
    "sp_adam_multi_peer": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), p_i64,
                                     C.POINTER(vp), C.POINTER(vp), p_i64, C.POINTER(vp),
                                     C.POINTER(vp), C.POINTER(vp), p_i64, C.POINTER(C.c_int32),
                                     f32, f32, f32, f32, vp]),
    "sp_sgd_multi": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp), p_i64, f32, vp]),
    "sp_tf32_split": (C.c_int, [vp, vp, i64, i64, C.c_int, vp]),
    "sp_tf32_round": (C.c_int, [vp, vp, i64, vp]),
    "sp_tf32_shadow_multi": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                       p_i64, p_i64, C.POINTER(C.c_int32), vp]),
    "sp_set_tf32_output_rounding": (C.c_int, [C.c_int]),
    "sp_conv2d_x3_plan": (C.c_int, [C.POINTER(ConvGeom), C.POINTER(C.c_int)]),
    "sp_nccl_load": (C.c_int, [C.c_char_p]),
    "sp_nccl_unique_id": (C.c_int, [vp]),
    "sp_nccl_init": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "sp_allreduce_sum_f32": (C.c_int, [vp, vp, i64, vp]),
    "sp_nccl_destroy": (C.c_int, [vp]),
    "sp_copy_multi": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp), p_i64, vp]),
    "sp_pack_multi": (C.c_int, [C.c_int, C.POINTER(vp), p_i64, p_i64, vp, vp]),
}

_lib = None

# ABI calls that enqueue at least one of our kernels (copies/memsets are not counted)
_LAUNCHING = {
    n for n in SIGNATURES
    if n not in ("sp_abi_version", "sp_last_error", "sp_device_info", "sp_debug_set_umma_policy",
                 "sp_debug_tma2_trace", "sp_debug_peer_trace", "sp_host_is_pinned", "sp_stream_create", "sp_event_create",
                 "sp_event_record", "sp_event_synchronize", "sp_host_pipe_threads", "sp_h2d_f64_as_f32_begin", "sp_h2d_f64_as_f32_finish", "sp_memcpy_d2h_async", "sp_stream_wait_event", "sp_memcpy_h2d", "sp_memcpy_d2h", "sp_memcpy_d2d", "sp_l2_flush",
                 "sp_conv2d_wgrad_workspace", "sp_set_tf32_output_rounding", "sp_conv2d_x3_plan",
                 "sp_set_scratch_lane", "sp_conv2d_wgrad_pooled_supported", "sp_conv2d_leaky_pool2_supported", "sp_conv2d_dgrad_unpool_supported", "sp_nccl_load", "sp_nccl_unique_id", "sp_nccl_init",
                 "sp_allreduce_sum_f32", "sp_nccl_destroy", "sp_peer_alloc", "sp_peer_open",
                 "sp_peer_close", "sp_peer_free", "sp_peer_exchange_create",
                 "sp_peer_exchange_destroy")
}


class SmallPebbleCudaError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """dlopen the in-tree library (building it with nvcc when absent) and type every export."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise SmallPebbleCudaError(f"{LIB_PATH} is missing: run `python -m smallpebble_b200.build`")
        from . import build as _build

        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.sp_abi_version() != 1:
        raise SmallPebbleCudaError("ABI version mismatch between _abi.py and the shared library")
    _lib = lib
    return lib


def last_error() -> str:
    return load().sp_last_error().decode(errors="replace")


class _Counter:
    """Kernel-launching ABI calls made by this process (bench: gpu_launches): the ones that
    went through ctypes plus the ones counted inside the fast-call binding."""

    __slots__ = ("_py",)

    def __init__(self):
        self._py = 0

    @property
    def launches(self):
        return self._py + (_fast.launches() if _fast is not None else 0)

    @launches.setter
    def launches(self, value):
        self._py = value - (_fast.launches() if _fast is not None else 0)


counter = _Counter()
_fast = None          # the _sp_fastcall module once loaded (None: ctypes only)
_fast_tried = False


def _load_fast():
    """Import smallpebble_b200/_lib/_sp_fastcall.so (csrc/fastcall.c); ctypes stays the
    fallback binding when it is missing or SP_NO_FASTCALL=1."""
    global _fast, _fast_tried
    if _fast_tried:
        return _fast
    _fast_tried = True
    if os.environ.get("SP_NO_FASTCALL") == "1":
        return None
    path = os.path.join(_HERE, "_lib", "_sp_fastcall.so")
    if not os.path.exists(path):
        return None
    try:
        import importlib.util

        spec = importlib.util.spec_from_file_location("_sp_fastcall", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _fast = mod
    except Exception:  # noqa: BLE001 - any failure means "use ctypes"
        _fast = None
    return _fast


def _fail(name, rc):
    raise SmallPebbleCudaError(f"{name} failed (code {rc}): {last_error()}")


class _Profiler:
    """Optional CUDA-event timing of kernel-launching ABI calls (bench.py's roofline).

    Two modes, both off by default (then an ABI call costs one attribute test):
      * `active = True`: every launching call gets an event pair (one instrumented step, to
        find the dominant call);  `seq` counts launching calls since `begin_step()`.
      * `arm(name, seq, every)`: ONLY calls of ABI function `name` pay anything: the one whose
        per-step sequence number is `seq` is bracketed by events on every `every`-th step.
        This is what runs inside the timed region."""

    def __init__(self):
        self._active = False
        self.select = None   # callable(name, seq) -> bool, or None (all)
        self.seq = 0
        self.step = 0
        self.records = []    # (name, seq, args, start_event, stop_event)
        self._armed = None   # (name, original wrapper)

    @property
    def active(self):
        return self._active

    @active.setter
    def active(self, value):
        # the per-call bookkeeping lives in the ctypes wrappers: swap bindings while it is on
        self._active = bool(value)
        f._rebind(slow=self._active)

    def begin_step(self):
        self.seq = 0
        self.step += 1

    def arm(self, name, seq, every=1):
        """Time call number `seq` (as counted in `active` mode) of function `name`."""
        self.disarm()
        inner = getattr(f, name)
        # position of the target among the calls of `name` within one step
        nth = sum(1 for r in self.records if r[0] == name and r[1] < seq)
        state = {"count": 0, "step": -1}
        prof_self = self

        def timed(*args):
            if state["step"] != prof_self.step:
                state["step"] = prof_self.step
                state["count"] = 0
            k = state["count"]
            state["count"] = k + 1
            if k == nth and prof_self.step % every == 0:
                import torch

                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                inner(*args)
                e1.record()
                prof_self.timed_records.append((name, seq, args, e0, e1))
            else:
                inner(*args)

        self.timed_records = []
        self._armed = (name, inner)
        setattr(f, name, timed)

    def disarm(self):
        if self._armed is not None:
            setattr(f, self._armed[0], self._armed[1])
            self._armed = None


prof = _Profiler()


def _checked(name, fn, launching):
    def fail(rc):
        raise SmallPebbleCudaError(f"{name} failed (code {rc}): {last_error()}")

    if launching:
        def wrapped(*args):
            if prof.active:
                seq = prof.seq
                prof.seq = seq + 1
                if prof.select is None or prof.select(name, seq):
                    import torch

                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rc = fn(*args)
                    e1.record()
                    if rc:
                        fail(rc)
                    counter.launches += 1
                    prof.records.append((name, seq, args, e0, e1))
                    return
            rc = fn(*args)
            if rc:
                fail(rc)
            counter.launches += 1
    else:
        def wrapped(*args):
            rc = fn(*args)
            if rc:
                fail(rc)
    wrapped.__name__ = name
    return wrapped


class _Functions:
    """Attribute access to status-checked exports: `abi.f.sp_gemm(...)`.  Loading is lazy so
    that importing the package works on a machine without the library (host-only tests);
    the first real call loads it or raises."""

    _slow_mode = False

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        lib = load()
        if name not in SIGNATURES:
            raise AttributeError(name)
        raw = getattr(lib, name)
        restype, argtypes = SIGNATURES[name]
        if restype is C.c_int and name != "sp_abi_version":
            fast = _load_fast()
            if fast is not None and not self._slow_mode:
                sig = "".join("f" if t is f32 else "p" for t in argtypes)
                fn = fast.bind(C.cast(raw, C.c_void_p).value, sig, name, _fail, name in _LAUNCHING)
            else:
                fn = _checked(name, raw, name in _LAUNCHING)
        else:
            fn = raw
        setattr(self, name, fn)
        return fn

    def _rebind(self, slow: bool):
        """Drop the cached bindings so that the next use re-creates them in the other mode."""
        if slow == self._slow_mode:
            return
        self._slow_mode = slow
        for name in [n for n in self.__dict__ if n in SIGNATURES]:
            delattr(self, name)


f = _Functions()


def dims(values):
    """int64[n] ctypes array for shapes / strides."""
    return (C.c_int64 * len(values))(*values)
