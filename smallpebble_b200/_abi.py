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
    "sp_memcpy_h2d": (C.c_int, [vp, vp, sz, vp]),
    "sp_memcpy_d2h": (C.c_int, [vp, vp, sz, vp]),
    "sp_memcpy_d2d": (C.c_int, [vp, vp, sz, vp]),
    "sp_fill_f32": (C.c_int, [vp, i64, f32, vp]),
    "sp_cast_f64_f32": (C.c_int, [vp, vp, i64, vp]),
    "sp_cast_i64_i32": (C.c_int, [vp, vp, i64, vp]),
    "sp_l2_flush": (C.c_int, [vp, sz, vp]),
    "sp_ewise_unary": (C.c_int, [C.c_int, vp, vp, i64, f32, vp]),
    "sp_ewise_binary": (C.c_int, [C.c_int, vp, p_i64, vp, p_i64, vp, C.c_int, p_i64, vp]),
    "sp_accumulate": (C.c_int, [vp, vp, i64, vp]),
    "sp_leaky_relu_fwd": (C.c_int, [vp, vp, i64, f32, vp]),
    "sp_leaky_relu_bwd": (C.c_int, [vp, vp, vp, i64, f32, vp]),
    "sp_where": (C.c_int, [vp, p_i64, vp, p_i64, vp, p_i64, vp, C.c_int, p_i64, vp]),
    "sp_strided_copy": (C.c_int, [vp, p_i64, vp, p_i64, C.c_int, p_i64, vp]),
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
    "sp_maxpool2d_fwd": (C.c_int, [vp, vp, vp] + [C.c_int] * 12 + [vp]),
    "sp_maxpool2d_bwd": (C.c_int, [vp, vp, vp] + [C.c_int] * 12 + [vp]),
    "sp_gemm": (C.c_int, [C.c_int, C.c_int, i64, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64,
                          i64, i64, C.c_int, vp]),
    "sp_conv2d_fprop": (C.c_int, [vp, vp, vp, C.POINTER(ConvGeom), C.c_int, vp]),
    "sp_conv2d_dgrad": (C.c_int, [vp, vp, vp, C.POINTER(ConvGeom), C.c_int, vp]),
    "sp_conv2d_wgrad_workspace": (sz, [C.POINTER(ConvGeom), C.c_int]),
    "sp_conv2d_wgrad": (C.c_int, [vp, vp, vp, C.POINTER(ConvGeom), C.c_int, vp, sz, vp]),
    "sp_adam_multi": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                C.POINTER(vp), p_i64, C.POINTER(vp), f32, f32, f32, f32, vp]),
    "sp_sgd_multi": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp), p_i64, f32, vp]),
    "sp_pack_multi": (C.c_int, [C.c_int, C.POINTER(vp), p_i64, p_i64, vp, vp]),
}

_lib = None

# ABI calls that enqueue at least one of our kernels (copies/memsets are not counted)
_LAUNCHING = {
    n for n in SIGNATURES
    if n not in ("sp_abi_version", "sp_last_error", "sp_device_info", "sp_debug_set_umma_policy",
                 "sp_debug_tma2_trace", "sp_memcpy_h2d", "sp_memcpy_d2h", "sp_memcpy_d2d", "sp_l2_flush",
                 "sp_conv2d_wgrad_workspace")
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
    __slots__ = ("launches",)

    def __init__(self):
        self.launches = 0


counter = _Counter()  # kernel-launching ABI calls made by this process (bench: gpu_launches)


class _Profiler:
    """Optional per-call CUDA-event timing of kernel-launching ABI calls (bench.py's roofline
    measurement).  `select(name, seq)` decides which calls get an event pair; `seq` counts
    launching calls since `begin_step()`.  Inactive by default: one attribute test per call."""

    def __init__(self):
        self.active = False
        self.select = None
        self.seq = 0
        self.records = []  # (name, seq, args, start_event, stop_event)

    def begin_step(self):
        self.seq = 0


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

    def __getattr__(self, name):
        lib = load()
        if name not in SIGNATURES:
            raise AttributeError(name)
        raw = getattr(lib, name)
        if SIGNATURES[name][0] is C.c_int and name != "sp_abi_version":
            fn = _checked(name, raw, name in _LAUNCHING)
        else:
            fn = raw
        setattr(self, name, fn)
        return fn


f = _Functions()


def dims(values):
    """int64[n] ctypes array for shapes / strides."""
    return (C.c_int64 * len(values))(*values)
