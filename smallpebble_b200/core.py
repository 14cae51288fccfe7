"""Variable, get_gradients and the ops -- the host-side mirror of the reference's operator
interface (same names, arguments, error behaviour as smallpebble/smallpebble.py, cited as
"sp.py:line"), with every array operation dispatched to the sm_100a kernels behind
include/smallpebble_b200.h.

Tape contract kept verbatim (sp.py:36-54): `Variable(array, local_gradients)` where
`local_gradients` is an iterable of `(child_variable, fn)` and `fn(path_value)` returns the
gradient contribution for the child.  User-defined ops written that way keep working; the
arrays they receive are `DeviceArray`s (NumPy functions accept them through `__array__`).

Differences from the reference, all internal:
  * `get_gradients` sweeps the tape in reverse topological order, calling each closure once
    with the summed upstream gradient.  The reference walks every root->leaf path
    (sp.py:79-83); both compute the same sums, ours without the 2x work softmax's diamond
    causes and without a full-tensor `+=` per path.
  * composite ops (conv2d, maxpool2d, softmax, cross_entropy) are single fused tape nodes:
    their internal nodes (padded image, patch view, im2col matrix, ...) never exist.
  * arrays are float32 on device (TF32 or fp32 math for contractions, `set_precision`).
"""

from __future__ import annotations

import ctypes as C
import functools
import itertools
import math
import operator
import os

import numpy as np

from . import _abi
from . import steady as _steady
from . import array as A
from .array import DeviceArray

F = _abi.f

# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------

# name -> (precision code of a single-pass contraction, error compensation on?)
_PRECISIONS = {"fp32": (_abi.PRECISION_FP32, False), "tf32": (_abi.PRECISION_TF32, True),
               "tf32x1": (_abi.PRECISION_TF32, False)}
_precision_name = os.environ.get("SP_PRECISION", "tf32").lower()
_precision, _compensate = _PRECISIONS[_precision_name]


# Library-global switch "round what the gradient kernels store to TF32" (csrc/common.cuh),
# mirrored here so that the ABI is only called when it changes -- in a training step it never
# does: see _want_rounding().
_rounding_state = False


def _apply_rounding_mode():
    F.sp_set_tf32_output_rounding(1 if _rounding_state else 0)


A.on_device_ready(_apply_rounding_mode)


def _reset_rounding() -> None:
    """Back to exact stores (used between tests that call the ABI directly)."""
    global _rounding_state
    if _rounding_state:
        _rounding_state = False
        if A.device_ready():
            _apply_rounding_mode()


def _want_rounding(child) -> bool:
    """Called by a gradient closure right before it launches the kernel that produces the
    gradient of `child`: when that gradient is going to be an operand of a tensor-core
    contraction (`child` is the output of conv2d / matmul, possibly through `+ bias` or a
    reshape) and the mode compensates, the kernel rounds what it stores to the nearest TF32
    value -- the tensor core would otherwise TRUNCATE it, a coherent bias.  Gradients that feed
    anything else are stored exactly."""
    global _rounding_state
    want = _compensate and getattr(child, "_to_contraction", False)
    if want != _rounding_state:
        _rounding_state = want
        _apply_rounding_mode()
    return want


def set_precision(name: str) -> None:
    """Math mode of the contractions (conv2d, matmul and their gradients).

    'tf32'   tcgen05 tensor cores, error-compensated where it matters: FORWARD contractions
             run as 3xTF32 (operands split into a TF32 value and its remainder, three products),
             so the discrete routing downstream -- leaky_relu's sign, maxpool2d's argmax --
             agrees with the float64 reference; backward contractions are single-pass with
             operands rounded to nearest instead of truncated.  Forward values ~1e-6, model
             gradients within 2e-3 of the float64 reference (csrc/tf32.cu).
    'tf32x1' every contraction a single tcgen05 pass on raw fp32 operands (the hardware
             truncates them to TF32): 2e-3 per op, but model-level gradients drift by a few
             percent because ~1e-4 of the routing decisions flip.  Fastest.
    'fp32'   CUDA-core FFMA (rel. tol 1e-5)."""
    global _precision, _compensate, _precision_name
    _precision, _compensate = _PRECISIONS[name.lower()]
    _precision_name = name.lower()
    _conv_plan.cache_clear()
    _reset_rounding()
    _steady.invalidate()


def get_precision() -> str:
    return _precision_name


def _fwd_native_precision() -> int:
    """Precision code of a FORWARD contraction for entry points that reach full accuracy by
    themselves (first-layer conv kernels, N <= 16 GEMMs): 3xTF32 when compensating."""
    return _abi.PRECISION_TF32X3 if _compensate else _precision


def _bwd_operand(g: DeviceArray) -> DeviceArray:
    """Operand of a single-pass BACKWARD contraction: rounded to the nearest TF32 values when
    compensating (most gradients arrive rounded by the kernel that produced them: no launch)."""
    if _compensate and not g._tf32:
        return A.tf32_round(g)
    return g


def debug_set_kernel_policy(policy: int) -> None:
    """Test hook (sp_debug_set_umma_policy): 0 default, 1 force the tcgen05 kernel, 2 force the
    CUDA-core kernel (+16: conservative producer synchronisation).  Cached conv plans hold
    policy-dependent workspace sizes, so they are dropped."""
    _abi.f.sp_debug_set_umma_policy(policy)
    _conv_plan.cache_clear()
    _steady.invalidate()


_fuse_softmax_xent = os.environ.get("SP_FUSE_SOFTMAX_XENT", "1") != "0"


_fuse_leaky_pool = os.environ.get("SP_FUSE_LEAKY_POOL", "1") != "0"
_fuse_bias = os.environ.get("SP_FUSE_BIAS", "1") != "0"
_fuse_conv_leaky = os.environ.get("SP_FUSE_CONV_LEAKY", "1") != "0"


def set_fusion(enabled: bool) -> None:
    """Fuse cross_entropy(softmax(x)) into one forward and one backward kernel (default on).
    With fusion the loss's tape edge goes straight to the logits, so `get_gradients` has no
    entry for the intermediate probability node; turn it off to get that entry."""
    global _fuse_softmax_xent, _fuse_leaky_pool, _fuse_bias, _fuse_conv_leaky
    _fuse_softmax_xent = bool(enabled)
    _fuse_leaky_pool = bool(enabled)
    _fuse_bias = bool(enabled)
    _fuse_conv_leaky = bool(enabled)
    _steady.invalidate()


# --------------------------------------------------------------------------------------------
# autodiff core (sp.py:36-87)
# --------------------------------------------------------------------------------------------


_serials = itertools.count()
_rebind_epoch = 0  # bumped by every `variable.array = ...` after construction
_serial_of = operator.attrgetter("_serial")


class Variable:
    """A value in a differentiable computation: `.array` (a DeviceArray) and
    `.local_gradients = [(child_variable, fn(path_value) -> array), ...]` (sp.py:36-54)."""

    def __init__(self, array, local_gradients=None) -> None:
        self._array = array if type(array) is DeviceArray else DeviceArray.from_host(array)
        self.local_gradients = local_gradients if local_gradients else []
        self._serial = next(_serials)  # construction order: get_gradients sorts the tape by it

    @property
    def array(self):
        return self._array

    @array.setter
    def array(self, value):
        global _rebind_epoch
        _rebind_epoch += 1  # optimisers cache parameter pointers until some array is rebound
        self._array = value if isinstance(value, DeviceArray) else DeviceArray.from_host(value)

    dtype = property(lambda self: self._array.dtype)  # sp.py:459-461
    ndim = property(lambda self: self._array.ndim)
    shape = property(lambda self: self._array.shape)


class Gradients(dict):
    """{Variable: DeviceArray} returned by get_gradients.  A plain dict, except that the
    contribution of a tape edge into a LEAF that is not flagged learnable (e.g. the gradient
    with respect to the input image batch) is evaluated the first time anybody asks for it:
    indexing, `in`, `get`, iteration, `len`, `items` ... all see the complete dictionary the
    reference returns (sp.py:87), but a training loop that only reads the learnables'
    gradients never pays for the image gradient.  SP_EAGER_LEAF_GRADS=1 disables deferral."""

    __slots__ = ("_deferred", "_steady", "_reduced")

    def __init__(self):
        super().__init__()
        self._deferred = {}
        self._steady = None  # steady.py: (recording, epoch) of a replayed backward pass
        self._reduced = False  # steady.py, data parallel: already summed over the ranks

    # plain dict access for get_gradients' own sweep (no deferred-leaf handling)
    get_plain = dict.get
    put_plain = dict.__setitem__

    def _defer(self, child, fn, path_value):
        self._deferred.setdefault(child, []).append((fn, path_value))

    def _materialise(self, key=None):
        if not self._deferred:
            return
        keys = [key] if key is not None else list(self._deferred)
        for k in keys:
            for fn, path_value in self._deferred.pop(k, ()):
                _accumulate(self, None, k, _as_device(fn(path_value)))

    def __missing__(self, key):
        if key in self._deferred:
            self._materialise(key)
            return dict.__getitem__(self, key)
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._deferred

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __iter__(self):
        self._materialise()
        return dict.__iter__(self)

    def __len__(self):
        return dict.__len__(self) + len(self._deferred)

    def keys(self):
        self._materialise()
        return dict.keys(self)

    def values(self):
        self._materialise()
        return dict.values(self)

    def items(self):
        self._materialise()
        return dict.items(self)


def _as_device(value) -> DeviceArray:
    return value if isinstance(value, DeviceArray) else DeviceArray.from_host(value)


def _accumulate(gradients, owned, child, value):
    have = dict.get(gradients, child)
    if have is None:
        dict.__setitem__(gradients, child, value)  # may alias the upstream buffer
    elif owned is not None and child in owned and have.shape == value.shape:
        A.accumulate_(have, value)
    else:
        dict.__setitem__(gradients, child, A.binary("add", have, value))
        if owned is not None:
            owned.add(child)


_eager_leaf_grads = os.environ.get("SP_EAGER_LEAF_GRADS", "0") == "1"
_dp_hook = None  # distributed.py: called after every write to a leaf's gradient


def _tape_order_dfs(variable):
    """Op nodes reachable from `variable`, parents before children (reverse post-order)."""
    order, seen, stack = [], {id(variable)}, [(variable, iter(variable.local_gradients))]
    while stack:
        node, it = stack[-1]
        advanced = False
        for child, _ in it:
            if id(child) not in seen:
                seen.add(id(child))
                stack.append((child, iter(child.local_gradients)))
                advanced = True
                break
        if not advanced:
            if node.local_gradients:
                order.append(node)
            stack.pop()
    order.reverse()
    return order


def _tape_order(variable):
    """Op nodes reachable from `variable`, parents before children.  Every Variable is numbered
    at construction and an op's result is constructed after its inputs, so descending serial
    numbers ARE a reverse topological order of the tape: collect the reachable op nodes with a
    plain stack and sort them (leaves have no edges to walk and are skipped)."""
    seen, stack, order = {id(variable)}, [variable], []
    while stack:
        node = stack.pop()
        edges = node.local_gradients
        if edges:
            order.append(node)
            for child, _ in edges:
                key = id(child)
                if key not in seen:
                    seen.add(key)
                    if child.local_gradients:  # leaves have nothing to walk
                        stack.append(child)
    try:
        order.sort(key=_serial_of, reverse=True)
    except AttributeError:  # a Variable subclass that skipped Variable.__init__
        order = _tape_order_dfs(variable)
    return order


_leaf_fork = None  # steady.py, while it records a backward pass: runs parameter-gradient
                   # closures on a second stream


def get_gradients(variable: Variable) -> dict:
    """d(sum of `variable`)/d(node) for every node reachable from `variable` (sp.py:57-87).

    Returns {Variable: DeviceArray}, keyed by identity, including `variable` itself (ones).
    """
    tag = getattr(variable, "_steady", None)
    if tag is not None:
        # the result of a replayed run (steady.py): its backward pass is a CUDA graph too
        replayed = _steady.gradients_of(variable, tag)
        if replayed is not None:
            return replayed
    return _get_gradients(variable)


def _get_gradients(variable: Variable) -> dict:
    order = _tape_order(variable)

    gradients = Gradients()
    seed_shape = variable.array.shape
    if math.prod(seed_shape) == 1:
        seed, fresh = A.scalar_one(seed_shape)  # shared read-only constant: no fill launch
    else:
        seed, fresh = DeviceArray.full(seed_shape, 1.0), True
    dict.__setitem__(gradients, variable, seed)
    # gradient buffers nobody else aliases: safe to accumulate in place
    owned = {variable} if fresh else set()
    get, put = gradients.get_plain, gradients.put_plain
    hook = _dp_hook
    fork = _leaf_fork
    for node in order:
        path_value = get(node)
        for child, multiply_by_locgrad in node.local_gradients:
            if not child.local_gradients:
                if getattr(child, "is_learnable", False):
                    if fork is not None:
                        # a parameter gradient feeds nothing else in this sweep: closure and
                        # accumulation run on the side stream, in order among themselves
                        def job(pv, child=child, fn=multiply_by_locgrad):
                            value = _as_device(fn(pv))
                            if get(child) is None:
                                put(child, value)
                            else:
                                _accumulate(gradients, owned, child, value)
                            return get(child)

                        job.child = child
                        job.fuses = getattr(multiply_by_locgrad, "fuses_pending", None)
                        fork(job, path_value)
                        continue
                elif not _eager_leaf_grads:
                    gradients._defer(child, multiply_by_locgrad, path_value)
                    continue
            value = multiply_by_locgrad(path_value)
            if type(value) is not DeviceArray:
                value = _as_device(value)
            if get(child) is None:
                put(child, value)  # first contribution: may alias the upstream buffer
            else:
                _accumulate(gradients, owned, child, value)
            if hook is not None and not child.local_gradients and getattr(child, "is_learnable", False):
                # data parallel: the exchange of the early gradients overlaps the rest of this
                # sweep (distributed._on_gradient_written)
                hook(child, get(child))
    return gradients


# --------------------------------------------------------------------------------------------
# shape / padding utilities (sp.py:659-782) -- host integer logic, bit-exact with the oracle
# --------------------------------------------------------------------------------------------


def broadcastinfo(a_shape, b_shape):
    "Get which dimensions are added or repeated when `a` and `b` are broadcast (sp.py:659-684)."
    return _broadcastinfo(tuple(a_shape), tuple(b_shape))


@functools.lru_cache(maxsize=1024)
def _broadcastinfo(a_shape, b_shape):
    ndim = max(len(a_shape), len(b_shape))
    pad_a, pad_b = ndim - len(a_shape), ndim - len(b_shape)
    ea = (1,) * pad_a + tuple(a_shape)
    eb = (1,) * pad_b + tuple(b_shape)
    if not all(p == q or p == 1 or q == 1 for p, q in zip(ea, eb)):
        raise ValueError(f"could not broadcast shapes {a_shape} {b_shape}")
    a_rep = tuple(i for i in range(ndim) if i < pad_a or (ea[i] == 1 and eb[i] > 1))
    b_rep = tuple(i for i in range(ndim) if i < pad_b or (eb[i] == 1 and ea[i] > 1))
    return a_rep, b_rep


def pad_amounts(array_length: int, stride: int, kern_length: int):
    """TF 'SAME' padding of a 1-D array (sp.py:721-730): the odd pixel goes to the end."""
    num_patches = math.ceil(array_length / stride)
    target_length = (num_patches - 1) * stride + kern_length
    padding_amount = max(target_length - array_length, 0)
    pad_front = padding_amount // 2
    return pad_front, padding_amount - pad_front


def _window_geometry(imheight, imwidth, kernheight, kernwidth, stride_y, stride_x, padding):
    """(pad_top, pad_bottom, pad_left, pad_right, outheight, outwidth): sp.py:733-761, 712-714."""
    if padding == "SAME":
        pad_top, pad_bottom = pad_amounts(imheight, stride_y, kernheight)
        pad_left, pad_right = pad_amounts(imwidth, stride_x, kernwidth)
    elif padding == "VALID":
        pad_top = pad_bottom = pad_left = pad_right = 0
    else:
        raise ValueError("padding must be 'SAME' or 'VALID'.")
    hp, wp = imheight + pad_top + pad_bottom, imwidth + pad_left + pad_right
    if hp < kernheight or wp < kernwidth:
        raise ValueError("window shape cannot be larger than input array shape")  # sp.py:710-711
    outheight = math.ceil((hp - kernheight + 1) / stride_y)
    outwidth = math.ceil((wp - kernwidth + 1) / stride_x)
    return pad_top, pad_bottom, pad_left, pad_right, outheight, outwidth


def padding2d(images, padding, imheight, imwidth, stride_y, stride_x, kernheight, kernwidth):
    "Pad `images` for conv2d, maxpool2d (sp.py:733-761).  Returns (images, imheight, imwidth)."
    if padding == "SAME":
        pad_top, pad_bottom = pad_amounts(imheight, stride_y, kernheight)
        pad_left, pad_right = pad_amounts(imwidth, stride_x, kernwidth)
        images = pad(images, pad_width=[(0, 0), (pad_top, pad_bottom), (pad_left, pad_right), (0, 0)])
        _, imheight, imwidth, _ = images.array.shape
    elif padding == "VALID":
        pass
    else:
        raise ValueError("padding must be 'SAME' or 'VALID'.")
    return images, imheight, imwidth


def np_strided_sliding_view(x: np.ndarray, window_shape, strides) -> np.ndarray:
    """Host helper kept for API parity (sp.py:687-718): strided window view of a NumPy array,
    shape = reduced + window.  The device ops use it only to build index maps."""
    if not len(window_shape) == x.ndim:
        raise ValueError("Must provide one window size for each dimension of x.")
    if not len(strides) == x.ndim:
        raise ValueError("Must provide one stride size for each dimension of x.")
    if any(size < 0 for size in window_shape):
        raise ValueError("`window_shape` cannot contain negative values")
    if any(stride < 0 for stride in strides):
        raise ValueError("`strides` cannot contain negative values")
    if any(x_size < w_size for x_size, w_size in zip(x.shape, window_shape)):
        raise ValueError("window shape cannot be larger than input array shape")
    reduced = tuple(math.ceil((n - w + 1) / s) for n, s, w in zip(x.shape, strides, window_shape))
    steps = tuple(b * s for b, s in zip(x.strides, strides))
    return np.lib.stride_tricks.as_strided(
        x, shape=reduced + tuple(window_shape), strides=steps + x.strides
    )


def patches_index(imheight, imwidth, kernheight, kernwidth, stride_y, stride_x):
    """im2col index map of one 2-D image (sp.py:764-782): ((rows, cols), outheight, outwidth,
    n_patches), rows/cols int64 [n_patches, kernheight*kernwidth]."""
    ys = np.arange(0, imheight - kernheight + 1, stride_y, dtype=np.int64)
    xs = np.arange(0, imwidth - kernwidth + 1, stride_x, dtype=np.int64)
    dy, dx = np.divmod(np.arange(kernheight * kernwidth, dtype=np.int64), kernwidth)
    rows = np.broadcast_to(ys[:, None, None] + dy[None, None, :], (len(ys), len(xs), len(dy)))
    cols = np.broadcast_to(xs[None, :, None] + dx[None, None, :], (len(ys), len(xs), len(dy)))
    outheight, outwidth = len(ys), len(xs)
    n_patches = outheight * outwidth
    return (
        (np.ascontiguousarray(rows).reshape(n_patches, -1),
         np.ascontiguousarray(cols).reshape(n_patches, -1)),
        outheight, outwidth, n_patches,
    )


def _unbroadcast(g: DeviceArray, shape, repeat_axes) -> DeviceArray:
    """Sum `g` back to an operand's own shape (the reference's enable_broadcast closures,
    sp.py:183-189) -- skipped entirely when nothing was broadcast."""
    if g.shape == tuple(shape):
        return g
    aux = g._aux
    if aux is not None and aux[0] == "colsum0" and repeat_axes == (0,) and g.ndim == 2:
        # the kernel that produced `g` already summed it over its rows (bias gradient)
        return aux[1].reshape(shape)
    return A.reduce_sum(g, repeat_axes).reshape(shape) if repeat_axes else g.reshape(shape)


# --------------------------------------------------------------------------------------------
# elementwise ops
# --------------------------------------------------------------------------------------------


def add(a: Variable, b: Variable) -> Variable:
    "Elementwise addition (sp.py:100-108)."
    ra, rb = broadcastinfo(a.array.shape, b.array.shape)
    sa, sb = a.array.shape, b.array.shape
    tag = a.array.pending_tag
    if (tag is not None and tag[0] == "matmul" and len(sa) == 2 and b.array.dtype == A.FLOAT
            and (sb == sa[1:] or sb == (1,) + sa[1:])):
        # the left operand is a layer matmul that has not been launched and the right one a
        # row vector: matmul + bias in one call (sp.py:630-635); the matmul node's own value
        # stays deferred unless somebody reads it
        av, bv, rows, n, k = tag[1:6]
        bias = b.array
        fwd_prec = tag[6] if len(tag) > 6 else _fwd_native_precision()

        def launch_gemm_bias(out):
            F.sp_gemm_bias(rows, n, k, av.ptr, bv.ptr, bias.ptr, out.ptr, fwd_prec, A.stream())

        if (_fuse_head and n <= 16 and rows * n) or (_fuse_linear and n > 16 and rows * n):
            # still not launched: softmax + cross_entropy may take the whole classifier head
            # into one kernel (core.cross_entropy), a leaky_relu the hidden layer with its
            # activation (core.leaky_relu); anything else launches it here
            value = DeviceArray.deferred(sa, launch_gemm_bias,
                                         ("gemm_bias", av, bv, bias, rows, n, k, fwd_prec), rows * n)
        else:
            value = DeviceArray.empty(sa)
            launch_gemm_bias(value)
    else:
        value = A.binary("add", a.array, b.array)
    out = Variable(value, [
        (a, lambda g: _unbroadcast(g, sa, ra)),
        (b, lambda g: _unbroadcast(g, sb, rb)),
    ])
    if sa == value.shape and getattr(a, "_to_contraction", False):
        out._to_contraction = True  # `matmul(x, w) + bias`: the gradient passes through to a
    return out


def sub(a: Variable, b: Variable) -> Variable:
    "Elementwise subtraction (sp.py:393-401)."
    ra, rb = broadcastinfo(a.array.shape, b.array.shape)
    value = A.binary("sub", a.array, b.array)
    sa, sb = a.array.shape, b.array.shape
    return Variable(value, [
        (a, lambda g: _unbroadcast(g, sa, ra)),
        (b, lambda g: A.unary("neg", _unbroadcast(g, sb, rb))),
    ])


def mul(a: Variable, b: Variable) -> Variable:
    "Elementwise multiplication (sp.py:303-311)."
    ra, rb = broadcastinfo(a.array.shape, b.array.shape)
    av, bv = a.array, b.array
    return Variable(A.binary("mul", av, bv), [
        (a, lambda g: _unbroadcast(A.binary("mul", g, bv), av.shape, ra)),
        (b, lambda g: _unbroadcast(A.binary("mul", g, av), bv.shape, rb)),
    ])


def div(a: Variable, b: Variable) -> Variable:
    "Elementwise division (sp.py:166-174)."
    ra, rb = broadcastinfo(a.array.shape, b.array.shape)
    av, bv = a.array, b.array
    return Variable(A.binary("div", av, bv), [
        (a, lambda g: _unbroadcast(A.binary("div", g, bv), av.shape, ra)),
        # -g * a / b^2
        (b, lambda g: _unbroadcast(A.binary("mul", g, A.binary("div_grad_b", av, bv)), bv.shape, rb)),
    ])


def neg(a: Variable) -> Variable:
    "Negate a variable (sp.py:314-318)."
    return Variable(A.unary("neg", a.array), [(a, lambda g: A.unary("neg", g))])


def exp(a: Variable) -> Variable:
    "Elementwise exp of `a` (sp.py:196-200)."
    value = A.unary("exp", a.array)
    return Variable(value, [(a, lambda g: A.binary("mul", g, value))])


def log(a: Variable) -> Variable:
    "Elementwise log of `a` (sp.py:224-228)."
    av = a.array
    return Variable(A.unary("log", av), [(a, lambda g: A.binary("div", g, av))])


def square(a: Variable) -> Variable:
    "Square each element of `a` (sp.py:367-375)."
    av = a.array
    return Variable(A.unary("square", av),
                    [(a, lambda g: A.binary("mul", g, A.unary("scale", av, 2.0)))])


_dgrad_unpool_cache = {}


def _dgrad_unpool_ok(cgeom, up_h, up_w, prec) -> bool:
    """Does the library scatter this conv's data gradient through a 2x2 max-pool in its own
    epilogue (sp_conv2d_dgrad_unpool_supported)?  Asked once per geometry."""
    key = (cgeom.N, cgeom.H, cgeom.W, cgeom.C, cgeom.KH, cgeom.KW, cgeom.K, cgeom.sy, cgeom.sx,
           cgeom.pad_t, cgeom.pad_l, cgeom.OH, cgeom.OW, up_h, up_w, prec, _steady.config_epoch)
    ok = _dgrad_unpool_cache.get(key)
    if ok is None:
        flag = C.c_int(0)
        _abi.f.sp_conv2d_dgrad_unpool_supported(cgeom, up_h, up_w, prec, flag)
        if len(_dgrad_unpool_cache) > 256:
            _dgrad_unpool_cache.clear()
        ok = _dgrad_unpool_cache[key] = bool(flag.value)
    return ok


def leaky_relu(a: Variable, alpha: float = 0.02) -> Variable:
    """Elementwise leaky relu (sp.py:231-236): x > 0 ? x : alpha*x; the slope at 0 is alpha.

    When the input is itself an op result (an activation, not a leaf that an optimiser may
    update in place) the launch is DEFERRED: the output is a pending array that a following
    `maxpool2d` consumes without it ever being stored (one fused kernel), and the gradient of
    that pool is folded into this op's gradient the same way.  Anything else that touches the
    output launches the ordinary kernel first, so results are bit-identical either way."""
    x = a.array
    alpha = float(alpha)

    def launch(out):
        tag = x.pending_tag
        if tag is not None and tag[0] == "conv2d" and alpha > 0:
            # the input is a conv2d that has not been launched: the activation is applied in
            # the conv kernel's epilogue and the pre-activation tensor is never written
            tag[1](out, alpha)
        elif tag is not None and tag[0] == "gemm_bias" and alpha > 0:
            # ... or a linear layer `matmul(x, w) + b` that has not been launched: one call
            _, gav, gbv, gbias, m, n, k, gprec = tag
            F.sp_gemm_bias_leaky(m, n, k, gav.ptr, gbv.ptr, gbias.ptr, out.ptr, gprec, alpha,
                                 A.stream())
        else:
            F.sp_leaky_relu_fwd(x.ptr, out.ptr, x.size, alpha, A.stream())

    if _fuse_leaky_pool and a.local_gradients and x.ndim == 4 and x.size:
        y = DeviceArray.deferred(x.shape, launch, ("leaky_relu", x, alpha), x.size)
    else:
        y = DeviceArray.empty_like(x)
        launch(y)

    def multiply_by_locgrad(g):
        dx = DeviceArray.empty_like(x)
        rounded = _want_rounding(a)
        tag = g.pending_tag if isinstance(g, DeviceArray) else None
        if tag is not None and tag[0] == "maxpool2d_bwd" and tag[1] is y:
            # the upstream gradient is a pool gradient that has not been launched: one kernel
            # scatters it and applies this op's slope
            _, _, g0, idx, geom, ypool = tag

            def launch_pool_bwd(out):
                r = _want_rounding(a)  # (the library-global switch may have moved since)
                if (x.pending_tag is not None and alpha > 0 and geom[4:8] == (2, 2, 2, 2)
                        and x.shape[3] % 4 == 0):
                    # the pool's input was never stored (conv + leaky_relu + pool ran as one
                    # kernel): the slope at the argmax is the slope at the pooled value
                    F.sp_maxpool2d_leaky_bwd_y(g0.ptr, idx.ptr, ypool.ptr, out.ptr, *geom, alpha,
                                               A.stream())
                else:
                    F.sp_maxpool2d_leaky_bwd(g0.ptr, idx.ptr, x.ptr, out.ptr, *geom, alpha,
                                             A.stream())
                out._tf32 = r  # rounded as stored: an operand of the conv gradients

            if alpha > 0 and getattr(a, "_pooled_wgrad", None) == geom:
                # `a` is a first-layer conv2d whose filter gradient can expand the POOLED
                # gradient itself (sp_conv2d_wgrad_pooled): not launched unless somebody else
                # needs the full-resolution gradient
                return DeviceArray.deferred(x.shape, launch_pool_bwd,
                                            ("pool_leaky_bwd", g0, idx, ypool, alpha), x.size)
            gtag = g0.pending_tag if type(g0) is DeviceArray else None
            if (gtag is not None and gtag[0] == "conv2d_dgrad" and alpha > 0 and _fuse_dgrad_unpool
                    and geom[4:10] == (2, 2, 2, 2, 0, 0)):
                # ... and the pooled gradient is itself a conv data gradient that has not been
                # launched: the conv kernel scatters through the argmax as it stores
                # (sp_conv2d_dgrad_unpool_leaky); the pooled gradient stays unlaunched
                _, cg, cw, cgeom, cprec = gtag
                if _dgrad_unpool_ok(cgeom, geom[1], geom[2], cprec):
                    rounded = _want_rounding(a)
                    F.sp_conv2d_dgrad_unpool_leaky(cg.ptr, cw.ptr, idx.ptr, ypool.ptr, dx.ptr, cgeom,
                                                   geom[1], geom[2], cprec, alpha, A.stream())
                    dx._tf32 = rounded
                    return dx
            launch_pool_bwd(dx)
            return dx
        # the slope depends on the sign of the input only; when the input was never stored
        # (conv + leaky ran as one kernel, alpha > 0) the output has the same sign
        src = y if (x.pending_tag is not None and y.pending_tag is None and alpha > 0) else x
        if tag is not None and tag[0] == "conv2d_dgrad":
            # the upstream gradient is a conv data gradient that has not been launched: the
            # conv kernel multiplies by this op's slope as it stores
            _, g0, cw, cgeom, cprec = tag
            F.sp_conv2d_dgrad_leaky(g0.ptr, cw.ptr, src.ptr, dx.ptr, cgeom, cprec, alpha, A.stream())
            dx._tf32 = rounded
            return dx
        g = _as_device(g)
        F.sp_leaky_relu_bwd(g.ptr, src.ptr, dx.ptr, x.size, alpha, A.stream())
        dx._tf32 = rounded
        return dx

    return Variable(y, [(a, multiply_by_locgrad)])


def where(condition: np.ndarray, a: Variable, b: Variable) -> Variable:
    "Condition is a boolean NumPy array, a and b are Variables (sp.py:418-441)."
    condition = np.asarray(condition, dtype=np.bool_)
    ra, rb = broadcastinfo(a.array.shape, b.array.shape)
    av, bv = a.array, b.array
    value = A.where(condition, av, bv)

    def grad_a(g):
        return _unbroadcast(A.where(condition, g, DeviceArray.zeros(())), av.shape, ra)

    def grad_b(g):
        return _unbroadcast(A.where(condition, DeviceArray.zeros(()), g), bv.shape, rb)

    return Variable(value, [(a, grad_a), (b, grad_b)])


# --------------------------------------------------------------------------------------------
# shape ops
# --------------------------------------------------------------------------------------------


def reshape(a: Variable, shape) -> Variable:
    "Reshape `a` into shape `shape` (sp.py:333-337): metadata only on device."
    src_shape = a.array.shape
    out = Variable(a.array.reshape(shape), [(a, lambda g: g.reshape(src_shape))])
    if getattr(a, "_to_contraction", False):
        out._to_contraction = True
    return out


def expand_dims(a: Variable, axis) -> Variable:
    "Add new axes with size of 1, indices specified by `axis` (sp.py:203-207)."
    src_shape = a.array.shape
    new_shape = np.expand_dims(np.empty(src_shape, np.bool_), axis).shape
    return Variable(a.array.reshape(new_shape), [(a, lambda g: g.reshape(src_shape))])


def matrix_transpose(a: Variable) -> Variable:
    "Swap the end two axes (sp.py:250-254)."
    return Variable(A.swapaxes(a.array, -2, -1), [(a, lambda g: A.swapaxes(g, -2, -1))])


def pad(a: Variable, pad_width) -> Variable:
    "Zero pad `a`, where pad_width[dim] = (left width, right width) (sp.py:321-330)."
    return Variable(A.pad(a.array, pad_width), [(a, lambda g: A.unpad(g, pad_width))])


def sum(a: Variable, axis=None) -> Variable:  # noqa: A001 (shadows builtin like the reference)
    "Sum elements of `a`, along axes specified in `axis` (sp.py:404-415)."
    src_shape = a.array.shape
    value = A.reduce_sum(a.array, axis)
    axes = A._norm_axes(axis, len(src_shape))
    keep_shape = tuple(1 if i in axes else n for i, n in enumerate(src_shape))

    def multiply_by_locgrad(g):
        # broadcast the reduced gradient back over the summed axes
        return A.binary("add", DeviceArray.zeros(src_shape), g.reshape(keep_shape))

    return Variable(value, [(a, multiply_by_locgrad)])


def getitem(a: Variable, indices) -> Variable:
    """Get elements of `a` using NumPy indexing (sp.py:210-221).  Basic indices (ints, slices,
    Ellipsis, None) are strided copies straight from the view geometry, forward and backward;
    advanced indices go through a flat index map (gather / scatter-add)."""
    src_shape = a.array.shape
    view = A.basic_view(src_shape, indices)
    if view is not None:
        value = A.view_gather(a.array, view)

        def scatter_view(g):
            result = DeviceArray.zeros(src_shape)
            A.view_scatter(result, view, _as_device(g))  # a basic view names no element twice
            return result

        return Variable(value, [(a, scatter_view)])
    flat = A.flat_index(src_shape, indices)
    value = A.gather_flat(a.array, flat)

    def multiply_by_locgrad(g):
        "(Takes into account elements indexed multiple times.)"
        result = DeviceArray.zeros(src_shape)
        A.scatter_add_flat(result, flat, g)
        return result

    return Variable(value, [(a, multiply_by_locgrad)])


def _pick_like(g, b_shape, picked):
    if picked.shape == b_shape:
        return picked
    return _unbroadcast(picked, b_shape, broadcastinfo(b_shape, picked.shape)[0])


def add_at(a: Variable, indices, b: Variable) -> Variable:
    """Add the elements of `b` to the locations in `a` specified by `indices`; repeated
    indices accumulate (sp.py:111-121)."""
    b_shape = b.array.shape
    value = a.array.copy()
    view = A.basic_view(a.array.shape, indices)
    if view is not None:
        A.view_scatter(value, view, b.array, accumulate=True)
        return Variable(value, [(a, lambda g: g),
                                (b, lambda g: _pick_like(g, b_shape, A.view_gather(_as_device(g), view)))])
    flat = A.flat_index(a.array.shape, indices)
    A.scatter_add_flat(value, flat, b.array)
    return Variable(value, [(a, lambda g: g),
                            (b, lambda g: _pick_like(g, b_shape, A.gather_flat(g, flat)))])


def setat(a: Variable, indices, b: Variable) -> Variable:
    """Similar to NumPy's `setitem`: a copy of `a` with `a[indices] = b` (sp.py:340-354)."""
    b_shape = b.array.shape
    value = a.array.copy()
    view = A.basic_view(a.array.shape, indices)
    if view is not None:
        A.view_scatter(value, view, b.array)

        def grad_a_view(g):
            out = _as_device(g).copy()  # the reference zeroes path_value in place; we do not
            A.view_scatter(out, view, DeviceArray.zeros(()))
            return out

        return Variable(value, [(a, grad_a_view),
                                (b, lambda g: _pick_like(g, b_shape, A.view_gather(_as_device(g), view)))])
    flat = A.flat_index(a.array.shape, indices)
    A.scatter_set_flat(value, flat, b.array)

    def grad_a(g):
        out = g.copy()  # the reference zeroes path_value in place; we leave it untouched
        A.scatter_set_flat(out, flat, DeviceArray.zeros(()))
        return out

    return Variable(value, [(a, grad_a),
                            (b, lambda g: _pick_like(g, b_shape, A.gather_flat(g, flat)))])


def strided_sliding_view(a: Variable, window_shape, strides) -> Variable:
    """Sliding window view with strides in index units (sp.py:378-390).  The view is described
    by strides alone -- positions x window, both walking the source -- so the device copies it
    out directly and the backward pass adds the gradient back through the same strides
    (np.add.at semantics: overlapping windows accumulate); no index map exists on either side."""
    src_shape = tuple(a.array.shape)
    window_shape, strides = tuple(window_shape), tuple(strides)
    # the reference's own checks and messages (sp.py:696-711)
    if len(window_shape) != len(src_shape):
        raise ValueError("Must provide one window size for each dimension of x.")
    if len(strides) != len(src_shape):
        raise ValueError("Must provide one stride size for each dimension of x.")
    if any(size < 0 for size in window_shape):
        raise ValueError("`window_shape` cannot contain negative values")
    if any(stride < 0 for stride in strides):
        raise ValueError("`strides` cannot contain negative values")
    if any(w > n for w, n in zip(window_shape, src_shape)):
        raise ValueError("window shape cannot be larger than input array shape")
    if 2 * len(src_shape) > _abi.MAX_DIMS:
        # more dimensions than the strided kernels take: gather through an index map
        index_map = np.array(np_strided_sliding_view(A._arange_index(src_shape), window_shape,
                                                     strides), order="C")
        value = A.gather_flat(a.array, index_map)

        def scatter_map(g):
            result = DeviceArray.zeros(src_shape)
            A.scatter_add_flat(result, index_map, g)
            return result

        return Variable(value, [(a, scatter_map)])
    cst = A._contig_strides(src_shape)
    positions = tuple((n - w) // s + 1 for n, w, s in zip(src_shape, window_shape, strides))
    view = (0, tuple(s * c for s, c in zip(strides, cst)) + tuple(cst), positions + window_shape)
    value = A.view_gather(a.array, view)

    def multiply_by_locgrad(g):
        result = DeviceArray.zeros(src_shape)
        A.view_scatter(result, view, _as_device(g), accumulate=True)
        return result

    return Variable(value, [(a, multiply_by_locgrad)])


def maxax(a: Variable, axis: int) -> Variable:
    "Reduce an axis, `axis`, to its max value, keeping it with size 1 (sp.py:257-279)."
    shape = a.array.shape
    nd = len(shape)
    axis = axis if axis >= 0 else nd + axis
    moved = a.array if axis == nd - 1 else A.swapaxes(a.array, axis, -1)
    n = moved.shape[-1]
    rows = moved.size // n if n else 0
    best = DeviceArray.empty((rows,))
    idx = DeviceArray.empty((rows,), np.int32)
    F.sp_argmax_rows_fwd(moved.ptr, best.ptr, idx.ptr, rows, n, A.stream())
    out_shape = tuple(1 if i == axis else v for i, v in enumerate(shape))
    moved_shape = moved.shape

    def multiply_by_locgrad(g):
        # the reference multiplies the upstream gradient, in its NATURAL layout, by the one-hot
        # swapped back to a's shape (sp.py:269-276): bring g into the row order of `moved`
        g = _as_device(g)
        gm = g if axis == nd - 1 else A.swapaxes(g.reshape(out_shape), axis, -1)
        dx = DeviceArray.empty(moved_shape)
        F.sp_argmax_rows_bwd(gm.ptr, idx.ptr, dx.ptr, rows, n, A.stream())
        return dx if axis == nd - 1 else A.swapaxes(dx, axis, -1)

    return Variable(best.reshape(out_shape), [(a, multiply_by_locgrad)])


# --------------------------------------------------------------------------------------------
# matmul (sp.py:239-247): NumPy batch broadcasting, gradients summed back over broadcast dims
# --------------------------------------------------------------------------------------------


def _gemm(ta, tb, m, n, k, a, lda, sa, b, ldb, sb, batch, out_shape, precision=None) -> DeviceArray:
    out = DeviceArray.empty(out_shape)
    if out.size:
        F.sp_gemm(ta, tb, m, n, k, a.ptr, lda, sa, b.ptr, ldb, sb, out.ptr, n, m * n, batch,
                  _precision if precision is None else precision, A.stream())
    return out


def _broadcast_to(x: DeviceArray, shape) -> DeviceArray:
    if x.shape == tuple(shape):
        return x
    return A.binary("add", DeviceArray.zeros(shape), x)


@functools.lru_cache(maxsize=1024)
def _matmul_plan(a_shape, b_shape):
    """Shape logic of a matmul call (validation, batch broadcasting), once per pair of shapes."""
    if len(a_shape) < 2 or len(b_shape) < 2:
        raise ValueError("matmul operands must have at least 2 dimensions")
    M, K = a_shape[-2:]
    K2, N = b_shape[-2:]
    if K != K2:
        raise ValueError(f"matmul: shapes {a_shape} and {b_shape} not aligned")
    a_batch, b_batch = a_shape[:-2], b_shape[:-2]
    ra, rb = broadcastinfo(a_batch, b_batch)
    batch_shape = tuple(np.broadcast_shapes(a_batch, b_batch))
    return M, K, N, a_batch, b_batch, ra, rb, batch_shape, math.prod(batch_shape)


_X3_MIN_FLOP = float(os.environ.get("SP_X3_MIN_FLOP", "2e8"))  # below: exact fp32 forward GEMM


def matmul(a: Variable, b: Variable) -> Variable:
    "Matrix multiplication (sp.py:239-247)."
    av, bv = a.array, b.array
    M, K, N, a_batch, b_batch, ra, rb, batch_shape, nb = _matmul_plan(av.shape, bv.shape)

    if not b_batch:
        # [..., M, K] x [K, N]: one GEMM over the flattened rows (the layer case)
        rows = math.prod(a_batch) * M
        if _fuse_bias and N <= 16 and a.local_gradients and len(a_batch) == 0 and rows and K:
            # the 10-way classifier: deferred so that `+ bias` (core.add) can run as part of
            # the same kernel; anything else that touches the product launches it as usual
            # (N <= 16 runs on the exact-fp32 skinny kernel: already fully accurate)
            fwd_prec = _fwd_native_precision()
            value = DeviceArray.deferred(
                (M, N),
                lambda out: F.sp_gemm(0, 0, rows, N, K, av.ptr, K, 0, bv.ptr, N, 0, out.ptr, N,
                                      rows * N, 1, fwd_prec, A.stream()),
                ("matmul", av, bv, rows, N, K))
        elif _compensate and N > 16 and rows and K and 2.0 * rows * N * K < _X3_MIN_FLOP:
            # a small product (the README MLP's layers: 31 MFLOP): exact fp32 on CUDA cores,
            # K split over the CTAs of a cluster -- one launch instead of operand splits plus a
            # tensor-core kernel that would run its 3K-long loop on two CTAs.  Deferred like
            # the classifier's: `+ bias` and a following leaky_relu join the same launch
            # (core.add, core.leaky_relu: sp_gemm_bias_leaky)
            if _fuse_linear and len(a_batch) == 0:
                value = DeviceArray.deferred(
                    (M, N),
                    lambda out: F.sp_gemm(0, 0, rows, N, K, av.ptr, K, 0, bv.ptr, N, 0, out.ptr, N,
                                          rows * N, 1, _abi.PRECISION_FP32, A.stream()),
                    ("matmul", av, bv, rows, N, K, _abi.PRECISION_FP32))
            else:
                value = _gemm(0, 0, rows, N, K, av, K, 0, bv, N, 0, 1, a_batch + (M, N),
                              _abi.PRECISION_FP32)
        elif _compensate and N > 16 and rows and K:
            # error-compensated forward product (3xTF32) as ONE tensor-core GEMM over a
            # contraction axis of 3K: [a_r | a_l | a_r] . [b_r ; b_r ; b_l]  (csrc/tf32.cu)
            a3 = A.tf32_split(av, _abi.SPLIT_RLR, K)
            b3 = A.tf32_split(bv, _abi.SPLIT_RRL, K * N)
            value = _gemm(0, 0, rows, N, 3 * K, a3, 3 * K, 0, b3, N, 0, 1, a_batch + (M, N))
        else:
            value = _gemm(0, 0, rows, N, K, av, K, 0, bv, N, 0, 1, a_batch + (M, N),
                          _fwd_native_precision())

        # N <= 16 (the classifier): all three products are tiny and run exactly -- the skinny
        # kernels are fp32 already, g . b^T contracts over N only -- so nothing is rounded
        exact = _compensate and N <= 16

        def grad_a(g):  # g . b^T
            if exact:
                return _gemm(0, 1, rows, K, N, g, N, 0, bv, N, 0, 1, av.shape, _abi.PRECISION_FP32)
            g, bw = _bwd_operand(g), _bwd_operand(bv)
            return _gemm(0, 1, rows, K, N, g, N, 0, bw, N, 0, 1, av.shape)

        def grad_b(g):  # a^T . g   (the sum over the batch is inside the contraction)
            if exact:
                return _gemm(1, 0, K, N, rows, av, K, 0, g, N, 0, 1, bv.shape, _fwd_native_precision())
            return _gemm(1, 0, K, N, rows, av, K, 0, _bwd_operand(g), N, 0, 1, bv.shape)

        out = Variable(value, [(a, grad_a), (b, grad_b)])
        out._to_contraction = not exact
        return out

    # general case: materialise broadcast batch dims, then strided-batched GEMMs
    # (compensated mode: the forward product takes the exact CUDA-core kernel)
    af = _broadcast_to(av, batch_shape + (M, K)) if a_batch != batch_shape else av
    bf = _broadcast_to(bv, batch_shape + (K, N)) if b_batch != batch_shape else bv
    value = _gemm(0, 0, M, N, K, af, K, M * K, bf, N, K * N, nb, batch_shape + (M, N),
                  _fwd_native_precision())

    def grad_a(g):
        full = _gemm(0, 1, M, K, N, _bwd_operand(g), N, M * N, _bwd_operand(bf), N, K * N, nb,
                     batch_shape + (M, K))
        return _unbroadcast(full, av.shape, ra)

    def grad_b(g):
        full = _gemm(1, 0, K, N, M, af, K, M * K, _bwd_operand(g), N, M * N, nb,
                     batch_shape + (K, N))
        return _unbroadcast(full, bv.shape, rb)

    out = Variable(value, [(a, grad_a), (b, grad_b)])
    out._to_contraction = True
    return out


# --------------------------------------------------------------------------------------------
# conv2d / maxpool2d
# --------------------------------------------------------------------------------------------


def _conv_geom(n, h, w, c, kh, kw, k, sy, sx, pt, pl, oh, ow) -> _abi.ConvGeom:
    return _abi.ConvGeom(n, h, w, c, kh, kw, k, sy, sx, pt, pl, oh, ow)


@functools.lru_cache(maxsize=1024)
def _conv_plan(x_shape, w_shape, padding, strides):
    """(ConvGeom, output shape, {precision: wgrad workspace bytes}) for one conv geometry."""
    n_images, imheight, imwidth, channels = x_shape
    kernheight, kernwidth, channels_in, channels_out = w_shape
    if channels != channels_in:
        raise ValueError(f"conv2d: images have {channels} channels, kernels expect {channels_in}")
    stride_y, stride_x = strides
    pt, _, pl, _, outh, outw = _window_geometry(
        imheight, imwidth, kernheight, kernwidth, stride_y, stride_x, padding)
    geom = _conv_geom(n_images, imheight, imwidth, channels, kernheight, kernwidth, channels_out,
                      stride_y, stride_x, pt, pl, outh, outw)
    out_shape = (n_images, outh, outw, channels_out)
    return geom, out_shape, {}, math.prod(out_shape), {}


def _conv_x3(plan):
    """How the forward contraction of this geometry runs error-compensated (csrc/tf32.cu):
    (X3 plan code, geometry the kernel sees, precision code, split kind / block of the images,
    split kind / block of the kernels) -- decided once per geometry by the library."""
    x3 = plan[4].get("x3")
    if x3 is None:
        g = plan[0]
        code = C.c_int(0)
        _abi.f.sp_conv2d_x3_plan(g, code)
        code = code.value
        if code == _abi.X3_NATIVE:
            x3 = (code, g, _abi.PRECISION_TF32X3, None, None)
        else:
            parts = 2 if code == _abi.X3_SPLIT2 else 3
            g2 = _conv_geom(g.N, g.H, g.W, parts * g.C, g.KH, g.KW, g.K, g.sy, g.sx, g.pad_t,
                            g.pad_l, g.OH, g.OW)
            if code == _abi.X3_SPLIT2:  # (r, l) interleaved per 32 channels, split-aware kernel
                x3 = (code, g2, _abi.PRECISION_TF32_SPLIT, (_abi.SPLIT_RL, 32),
                      (_abi.SPLIT_RL, 32 * g.K))
            else:                       # plain K-concatenation: [r | l | r] . [r ; r ; l]
                x3 = (code, g2, _abi.PRECISION_TF32, (_abi.SPLIT_RLR, g.C),
                      (_abi.SPLIT_RRL, g.C * g.K))
        plan[4]["x3"] = x3
    return x3


def conv2d(images: Variable, kernels: Variable, padding: str = "SAME", strides=(1, 1)) -> Variable:
    """2D convolution, with same api as tf.nn.conv2d (sp.py:124-163).

    images [n_images, imheight, imwidth, n_channels] (NHWC), kernels [kernheight, kernwidth,
    channels_in, channels_out] (HWIO) -> [n_images, outheight, outwidth, channels_out].
    One implicit-GEMM kernel; the gradients are two more (input gradient in gather form
    instead of np.add.at col2im, filter gradient with a deterministic split reduction).
    """
    x, w = images.array, kernels.array
    plan = _conv_plan(x.shape, w.shape, padding, tuple(strides))
    geom, out_shape = plan[0], plan[1]
    prec = _precision
    compensate = _compensate

    def launch(out, alpha=None):
        """y = conv2d(x, w), or leaky_relu(conv2d(x, w), alpha) in the kernel's epilogue."""
        if not out.size:
            return
        xa, wa, g, pr = x, w, geom, prec
        if compensate:
            # forward contraction error-compensated (3xTF32): the routing decisions taken on
            # this tensor (leaky_relu sign, maxpool argmax) must match the float64 reference
            _, g, pr, xsplit, wsplit = _conv_x3(plan)
            if xsplit is not None:
                if images.local_gradients and getattr(images, "_argmax", None) is not None:
                    _split_hints[x.shape] = xsplit  # (maxpool2d then produces it as it pools)
                xa = A.tf32_split(x, *xsplit)
                wa = A.tf32_split(w, *wsplit)
        if alpha is None:
            F.sp_conv2d_fprop(xa.ptr, wa.ptr, out.ptr, g, pr, A.stream())
        else:
            F.sp_conv2d_fprop_leaky(xa.ptr, wa.ptr, out.ptr, g, pr, alpha, A.stream())

    def launch_pooled(ypool, idx, split, alpha):
        """maxpool2d(leaky_relu(conv2d(x, w), alpha), 2, 2, strides [2, 2]) in ONE kernel
        (sp_conv2d_leaky_pool2_fwd: first-layer geometries) -- the full-resolution tensor stays
        unlaunched.  False: no such kernel for this conv, the caller pools the usual way."""
        xa, wa, g, pr = x, w, geom, prec
        xsplit = wsplit = None
        if compensate:
            code, g, pr, xsplit, wsplit = _conv_x3(plan)
            if xsplit is not None and code != _abi.X3_SPLIT2:
                return False  # (K-concatenated operands: not a layout the pooled epilogue takes)
        key = ("pool2", pr, _steady.config_epoch)
        ok = plan[4].get(key)
        if ok is None:
            flag = C.c_int(0)
            _abi.f.sp_conv2d_leaky_pool2_supported(g, pr, flag)
            ok = plan[4][key] = bool(flag.value)
        if not ok:
            return False
        if xsplit is not None:
            # error-compensated tcgen05 layer: pre-split (r, l) operands, as in launch()
            if images.local_gradients and getattr(images, "_argmax", None) is not None:
                _split_hints[x.shape] = xsplit
            xa = A.tf32_split(x, *xsplit)
            wa = A.tf32_split(w, *wsplit)
        F.sp_conv2d_leaky_pool2_fwd(xa.ptr, wa.ptr, ypool.ptr, idx.ptr,
                                    split._ptr if split is not None else None, g, pr, alpha,
                                    A.stream())
        return True

    out_size = plan[3]
    if _fuse_conv_leaky and out_size:
        # not launched yet: a leaky_relu consumer folds its activation into this kernel's
        # epilogue (core.leaky_relu), a leaky_relu + 2x2 maxpool2d consumer the pooling too
        # (core.maxpool2d); anything else that touches the result launches it as usual
        y = DeviceArray.deferred(out_shape, launch, ("conv2d", launch, launch_pooled), out_size)
    else:
        y = DeviceArray.empty(out_shape, A.FLOAT, out_size)
        launch(y)

    def grad_images(g):
        prec_b = _precision
        # single-pass contraction g . w^T: operands rounded to nearest when compensating (the
        # gradient usually arrives rounded by its producer, the weights from their shadow)
        g = _bwd_operand(_as_device(g))
        wr = _bwd_operand(w)

        def launch_dgrad(out):
            if out.size:
                F.sp_conv2d_dgrad(g.ptr, wr.ptr, out.ptr, geom, prec_b, A.stream())

        if _fuse_conv_leaky and x.size:
            # not launched yet: when the consumer is a leaky_relu's gradient closure, the slope
            # is applied as the data gradient is stored (sp_conv2d_dgrad_leaky)
            return DeviceArray.deferred(x.shape, launch_dgrad, ("conv2d_dgrad", g, wr, geom, prec_b),
                                        x.size)
        dx = DeviceArray.empty_like(x)
        launch_dgrad(dx)
        return dx

    def fuses_pending(g):
        "True when grad_kernels consumes this not-yet-launched gradient without launching it."
        tag = g.pending_tag if type(g) is DeviceArray else None
        return tag is not None and tag[0] == "pool_leaky_bwd" and pooled_geom is not None

    def grad_kernels(g):
        dw = DeviceArray.empty_like(w)
        ws_bytes = plan[2].get(_precision)
        if ws_bytes is None:
            ws_bytes = plan[2][_precision] = _abi.load().sp_conv2d_wgrad_workspace(geom, _precision)
        ws = DeviceArray.empty((ws_bytes // 4,)) if ws_bytes else None
        if fuses_pending(g):
            # the gradient is leaky_relu + 2x2 maxpool2d's, still pooled: the kernel scatters it
            # to the argmax positions as it stages its tiles (bit-identical to the two launches)
            _, g0, idx, ypool, alpha = g.pending_tag
            F.sp_conv2d_wgrad_pooled(x.ptr, g0.ptr, idx.ptr, ypool.ptr, dw.ptr, geom, alpha,
                                     _precision, ws.ptr if ws is not None else None, ws_bytes,
                                     A.stream())
            return dw
        g = _bwd_operand(_as_device(g))
        F.sp_conv2d_wgrad(x.ptr, g.ptr, dw.ptr, geom, _precision,
                          ws.ptr if ws is not None else None, ws_bytes, A.stream())
        return dw

    grad_kernels.fuses_pending = fuses_pending
    out = Variable(y, [(images, grad_images), (kernels, grad_kernels)])
    out._to_contraction = True  # gradients w.r.t. this tensor are operands of contractions
    # geometry of the 2x2 / stride-2 pool whose (leaky_relu'd) gradient the filter gradient can
    # take in pooled form (first-layer convs, TF32): core.leaky_relu then leaves it pooled
    pooled_geom = _pooled_wgrad_geom(plan, out_shape) if (_fuse_leaky_pool and _pooled_wgrad) else None
    if pooled_geom is not None:
        out._pooled_wgrad = pooled_geom
    return out


_pooled_wgrad = os.environ.get("SP_POOLED_WGRAD", "1") != "0"
_fuse_conv_pool = os.environ.get("SP_FUSE_CONV_POOL", "1") != "0"  # conv + leaky + pool, one kernel
_fuse_head = os.environ.get("SP_FUSE_HEAD", "1") != "0"  # matmul + bias + softmax + cross_entropy
_fuse_linear = os.environ.get("SP_FUSE_LINEAR", "1") != "0"  # small matmul + bias + leaky_relu
# conv data gradient scattered through the following pool / leaky_relu backward in the conv
# kernel's epilogue (sp_conv2d_dgrad_unpool_leaky).  OFF by default: alone the fused kernel beats
# the two launches (15.1 vs 18.2 us on the README CNN's conv3), but inside the backward graph the
# step gets 2.4 us LONGER -- the heavy one-CTA-per-SM conv kernels of the two streams queue for
# SMs, and the light pool kernel used to run beside them while the fused epilogue keeps 100 SMs
# 3.6 us longer (scripts/steady_parts.py).
_fuse_dgrad_unpool = os.environ.get("SP_FUSE_DGRAD_UNPOOL", "0") == "1"


def _pooled_wgrad_geom(plan, out_shape):
    key = ("pooled", _precision)
    cached = plan[4].get(key, 0)
    if cached == 0:
        cached = None
        n, oh, ow, k = out_shape
        if _precision == _abi.PRECISION_TF32 and oh % 2 == 0 and ow % 2 == 0 and oh and ow:
            flag = C.c_int(0)
            _abi.f.sp_conv2d_wgrad_pooled_supported(plan[0], _precision, flag)
            if flag.value:
                cached = (n, oh, ow, k, 2, 2, 2, 2, 0, 0, oh // 2, ow // 2)
        plan[4][key] = cached
    return cached


_UINT8 = np.dtype(np.uint8)
_split_hints = {}  # pooled output shape -> (split kind, block) its consumer conv2d asked for


@functools.lru_cache(maxsize=256)
def _pool_plan(x_shape, kernheight, kernwidth, stride_y, stride_x, padding):
    """(kernel geometry arguments, output shape) of a pooling call: host integer logic that
    is the same every step of a training loop."""
    n_images, imheight, imwidth, n_channels = x_shape
    pt, _, pl, _, outh, outw = _window_geometry(
        imheight, imwidth, kernheight, kernwidth, stride_y, stride_x, padding)
    geom = (n_images, imheight, imwidth, n_channels, kernheight, kernwidth, stride_y, stride_x,
            pt, pl, outh, outw)
    out_shape = (n_images, outh, outw, n_channels)
    return geom, out_shape, math.prod(out_shape)


def maxpool2d(images: Variable, kernheight: int, kernwidth: int, padding: str = "SAME",
              strides=[1, 1]) -> Variable:  # noqa: B006 (signature kept from sp.py:282-288)
    """Maxpooling on a `variable` of shape [n_images, imheight, imwidth, n_channels]
    (sp.py:282-300).  Zero padding takes part in the max and the first maximum in (dy, dx)
    order wins -- exactly what np.pad + np.argmax give the reference."""
    x = images.array
    stride_y, stride_x = strides
    geom, out_shape, out_size = _pool_plan(x.shape, kernheight, kernwidth, stride_y, stride_x,
                                           padding)
    y = DeviceArray.empty(out_shape, A.FLOAT, out_size)
    idx = DeviceArray.empty(out_shape, _UINT8, out_size)
    tag = x.pending_tag
    fused = tag is not None and tag[0] == "leaky_relu" and y.size > 0
    hint = _split_hints.get(out_shape) if _compensate and y.size else None
    conv_pool = None
    if fused and _fuse_conv_pool and tag[2] > 0 and geom[4:10] == (2, 2, 2, 2, 0, 0):
        ctag = tag[1].pending_tag
        if ctag is not None and ctag[0] == "conv2d" and len(ctag) > 2:
            conv_pool = ctag[2]  # the conv that feeds the leaky_relu has not been launched
    done = False
    if conv_pool is not None and (hint is None or hint == (_abi.SPLIT_RL, min(out_shape[3], 32))):
        # conv2d + leaky_relu + this pool as one kernel (first-layer convs, and the layers of
        # the 2-SM tcgen05 kernel, which pools in its epilogue); the (r, l) split of the pooled
        # tensor comes out of the same kernel (blocks of 32 channels, or whole pixels below that)
        split = DeviceArray.empty((out_size * 2,), A.FLOAT, out_size * 2) if hint is not None else None
        done = conv_pool(y, idx, split, tag[2])
        if done and split is not None:
            y._shadow = [A.inplace_epoch, None, {hint: split}]
    if done:
        pass
    elif hint is not None:
        # a compensated conv2d consumed the previous tensor of this shape as (r, l) split
        # operand: this kernel writes that form beside y (tf32.cu), no split launch later
        kind, block = hint
        parts = 2 if kind == _abi.SPLIT_RL else 3
        split = DeviceArray.empty((out_size * parts,), A.FLOAT, out_size * parts)
        src, leaky, alpha = (tag[1], 1, tag[2]) if fused else (x, 0, 0.0)
        F.sp_maxpool2d_fwd_split(src.ptr, y.ptr, idx.ptr, split._ptr, block, kind, *geom, leaky,
                                 alpha, A.stream())
        y._shadow = [A.inplace_epoch, None, {hint: split}]
    elif fused:
        # the input is a leaky_relu that has not been launched: pool leaky(src) directly
        F.sp_maxpool2d_leaky_fwd(tag[1].ptr, y.ptr, idx.ptr, *geom, tag[2], A.stream())
    elif y.size:
        F.sp_maxpool2d_fwd(x.ptr, y.ptr, idx.ptr, *geom, A.stream())

    def multiply_by_locgrad(g):
        g = _as_device(g)

        def launch(out):
            if out.size:
                rounded = _want_rounding(images)
                F.sp_maxpool2d_bwd(g.ptr, idx.ptr, out.ptr, *geom, A.stream())
                out._tf32 = rounded

        if fused and x.size:
            # deferred: leaky_relu's gradient closure folds this scatter into its own kernel
            return DeviceArray.deferred(x.shape, launch, ("maxpool2d_bwd", x, g, idx, geom, y),
                                        x.size)
        dx = DeviceArray.empty_like(x)
        launch(dx)
        return dx

    out = Variable(y, [(images, multiply_by_locgrad)])
    out._argmax = idx  # uint8 [N,OH,OW,C]: dy*kernwidth+dx, exposed for the bit-exact tests
    return out


# --------------------------------------------------------------------------------------------
# softmax / cross entropy
# --------------------------------------------------------------------------------------------


def softmax(a: Variable, axis: int = -1) -> Variable:
    """Softmax on `axis` (sp.py:357-364).  The kernel shifts by the per-slice maximum instead
    of the reference's whole-array maximum: identical mathematics (shift invariance), closest
    to the float64 result over the widest input range in fp32."""
    x = a.array
    nd = x.ndim
    ax = axis if axis >= 0 else nd + axis
    outer, n, inner = math.prod(x.shape[:ax]), x.shape[ax], math.prod(x.shape[ax + 1:])
    fusable = inner == 1 and nd == 2

    def launch(out):
        if out.size:
            F.sp_softmax_fwd(x.ptr, out.ptr, outer, n, inner, A.stream())

    if fusable and _fuse_softmax_xent and a.local_gradients and x.size:
        # Not launched yet: cross_entropy computes the probabilities and the loss from the
        # logits in ONE kernel and hands this array its buffer; anything else that touches the
        # probabilities first launches the ordinary kernel (bitwise the same values).
        y = DeviceArray.deferred(x.shape, launch, ("softmax", x), x.size)
    else:
        y = DeviceArray.empty_like(x)
        launch(y)

    def multiply_by_locgrad(g):
        dx = DeviceArray.empty_like(x)
        if dx.size:
            rounded = _want_rounding(a)
            F.sp_softmax_bwd(y.ptr, g.ptr, dx.ptr, outer, n, inner, A.stream())
            dx._tf32 = rounded and inner == 1
        return dx

    out = Variable(y, [(a, multiply_by_locgrad)])
    if fusable:
        out._softmax_of = a  # lets cross_entropy fuse forward and backward with this node
    return out


def cross_entropy(y_pred: Variable, y_true: np.ndarray, axis: int = -1) -> Variable:
    """Cross entropy (sp.py:612-621): y_pred [batch, n_classes] probabilities, y_true integer
    labels [batch]; returns the SUM over the batch as a 0-d Variable.

    When `y_pred` is the direct output of `sp.softmax` the loss is evaluated from the logits
    (logsumexp - logit) and the backward edge goes straight to the logits as
    g * (p - onehot): the fused softmax-cross-entropy of the README models."""
    probs = y_pred.array
    if probs.ndim != 2:
        raise ValueError("cross_entropy expects y_pred of shape [batch_size, n_classes]")
    rows, cols = probs.shape
    if isinstance(y_true, DeviceArray) and y_true.dtype == np.int32:
        labels = y_true  # already resident on the device (validated by whoever uploaded it)
        if labels.size != rows:
            raise ValueError("cross_entropy: len(y_true) must equal the batch size")
    else:
        labels_host = np.asarray(y_true.array if isinstance(y_true, Variable) else y_true)
        labels_host = labels_host.reshape(-1).astype(np.int64)
        if labels_host.shape[0] != rows:
            raise ValueError("cross_entropy: len(y_true) must equal the batch size")
        if rows and (labels_host.min() < -cols or labels_host.max() >= cols):
            raise IndexError("cross_entropy: label out of range")
        labels_host = np.where(labels_host < 0, labels_host + cols, labels_host)
        labels = DeviceArray.from_host(labels_host, dtype=np.int32)
    loss = DeviceArray.empty(())
    logits_var = getattr(y_pred, "_softmax_of", None) if _fuse_softmax_xent else None

    if logits_var is not None:
        logits = logits_var.array
        if probs.pending_tag is not None:
            # the softmax has not been launched: this kernel produces its output as well
            probs._adopt(DeviceArray.empty_like(probs))
        # (re)computes probs into the softmax node's buffer (bitwise the same values) + the loss
        pre = None
        if (_fuse_bias and 0 < rows <= 2048 and rows * cols <= 32768
                and not (_compensate and getattr(logits_var, "_to_contraction", False))):
            # ... and, in the same launch, the gradient of the logits for a unit seed with its
            # column sums: what grad_logits returns when this loss is the root of get_gradients
            pre = (DeviceArray.empty_like(logits), DeviceArray.empty((cols,)))
            ltag = logits.pending_tag
            if ltag is not None and ltag[0] == "gemm_bias" and _fuse_head:
                # ... and the classifier's matmul + bias, which has not been launched either:
                # the whole head is one kernel (the GEMM's last block runs the softmax)
                _, av, bv, bias, m, n, k, fwd_prec = ltag
                logits._adopt(DeviceArray.empty_like(logits))
                F.sp_gemm_bias_softmax_xent(m, n, k, av.ptr, bv.ptr, bias.ptr, labels.ptr,
                                            logits._ptr, probs.ptr, loss.ptr, pre[0]._ptr,
                                            pre[1]._ptr, fwd_prec, A.stream())
            else:
                F.sp_softmax_xent_fwd_grad(logits.ptr, labels.ptr, probs.ptr, loss.ptr,
                                           pre[0]._ptr, pre[1]._ptr, rows, cols, A.stream())
        else:
            F.sp_softmax_xent_fwd(logits.ptr, labels.ptr, probs.ptr, loss.ptr, rows, cols,
                                  A.stream())

        def grad_logits(g):
            if pre is not None and g._aux is A.ONES and not _want_rounding(logits_var):
                dlogits = pre[0]
                dlogits._aux = ("colsum0", pre[1])
                dlogits._tf32 = False
                return dlogits
            dlogits = DeviceArray.empty_like(logits)
            rounded = _want_rounding(logits_var)
            if dlogits.size and _fuse_bias and rows * cols <= 32768:
                # also emits the column sums: the gradient of a bias added to the logits, which
                # core._unbroadcast picks up instead of launching a reduction
                colsum = DeviceArray.empty((cols,))
                F.sp_softmax_xent_bwd_colsum(probs.ptr, labels.ptr, g.ptr, dlogits.ptr, colsum.ptr,
                                             rows, cols, A.stream())
                dlogits._aux = ("colsum0", colsum)
            elif dlogits.size:
                F.sp_softmax_xent_bwd(probs.ptr, labels.ptr, g.ptr, dlogits.ptr, rows, cols,
                                      A.stream())
            dlogits._tf32 = rounded  # rounded as stored: operand of the classifier's grads
            return dlogits

        out = Variable(loss, [(logits_var, grad_logits)])
        out._fused_softmax_xent = True
        return out

    F.sp_xent_fwd(probs.ptr, labels.ptr, loss.ptr, rows, cols, A.stream())

    def grad_probs(g):
        dprobs = DeviceArray.empty_like(probs)
        if dprobs.size:
            F.sp_xent_bwd(probs.ptr, labels.ptr, g.ptr, dprobs.ptr, rows, cols, A.stream())
        return dprobs

    return Variable(loss, [(y_pred, grad_probs)])


# --------------------------------------------------------------------------------------------
# operator overloads and attributes (sp.py:449-461)
# --------------------------------------------------------------------------------------------

Variable.__add__ = add
Variable.__mul__ = mul
Variable.__neg__ = neg
Variable.__sub__ = sub
Variable.__truediv__ = div
Variable.__getitem__ = getitem
