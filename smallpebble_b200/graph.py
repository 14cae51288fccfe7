"""Lazy graphs, learnables, optimisers and layer constructors: the reference's L3 layer
(sp.py:472-653) kept as host Python (BASELINE.json: "Host code stays Python").  Initialisers
use the global NumPy RNG exactly like the reference, so `np.random.seed(s)` gives the same
weights on both sides (cast once to fp32 on upload).  Adam / SGD run as one fused
multi-tensor kernel; under `smallpebble_b200.distributed` they first all-reduce the gradients.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from . import array as A
from . import core
from . import distributed as _distributed
from . import steady as _steady
from .array import DeviceArray
from .core import Variable

F = _abi.f


# ---------------- LAZY GRAPHS (sp.py:472-520)


class AssignmentError(Exception):
    pass


class Lazy:
    """A lazy node: `Lazy(fn)(*args)`, evaluated by `.run()` (sp.py:476-509).  No memoisation,
    like the reference: a node shared by two parents is evaluated once per parent."""

    def __init__(self, function) -> None:
        self.function = function
        self.arguments = []
        self._flag_args = None  # the (immutable) argument tuple `_flags` was computed for
        self._flags = ()

    def __call__(self, *args):
        if self.arguments:
            raise AssignmentError(f"Arguments already assigned to {self}.")
        self.arguments = args
        return self

    def run(self):
        if _steady.depth == 0 and _steady.enabled:
            # outermost call: a root whose steps have settled is replayed from CUDA graphs
            return _steady.run_root(self)
        return self._run()

    def _run(self):
        args = self.arguments
        if not args:
            raise AssignmentError(f"No arguments have been assigned to {self}.")
        if type(args) is tuple:
            # which arguments are lazy nodes is fixed for a tuple: decide once, not per run
            if args is not self._flag_args:
                self._flags = tuple(hasattr(a, "run") for a in args)
                self._flag_args = args
            return self.function(*[a.run() if lazy else a for a, lazy in zip(args, self._flags)])
        return self.function(*[a.run() if hasattr(a, "run") else a for a in args])


class Placeholder(Lazy):
    "A placeholder lazy graph node, in which to place SmallPebble variables (sp.py:512-520)."

    def __init__(self):
        super().__init__(lambda a: a)

    def assign_value(self, variable) -> None:
        self.arguments = [variable]

    def run(self):
        args = self.arguments
        if not args:
            raise AssignmentError(f"No arguments have been assigned to {self}.")
        if len(args) != 1:
            return super().run()
        a = args[0]
        return a.run() if hasattr(a, "run") else a


# ---------------- LEARNABLES (sp.py:526-543)


def learnable(variable: Variable) -> Variable:
    "Flag `variable` as learnable."
    variable.is_learnable = True
    return variable


def get_learnables(lazy_node: Lazy) -> list:
    "Get `variables` where is_learnable=True from a lazy graph (depth-first argument order)."
    learnable_vars = []

    def find_learnables(node):
        for child in getattr(node, "arguments", []):
            if getattr(child, "is_learnable", False):
                learnable_vars.append(child)
            find_learnables(child)

    find_learnables(lazy_node)
    return learnable_vars


# ---------------- OPTIMISERS (sp.py:552-583, 645-653)


def _ptr_table(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


def _device_gradient(variable: Variable, gradients) -> DeviceArray:
    g = gradients[variable]  # KeyError when missing, like the reference
    if not isinstance(g, DeviceArray):
        g = DeviceArray.from_host(g)
    if g.shape != variable.array.shape:
        g = A.binary("add", DeviceArray.zeros(variable.array.shape), g)
    return g


class _StepCounts:
    """`adam.t[variable]` (sp.py:569): completed steps of each variable, read from the
    device-resident counters the fused kernel maintains (so it stays right under CUDA-graph
    replay, where no Python runs per step)."""

    def __init__(self, counters):
        self._counters = counters

    def __getitem__(self, variable):
        if variable not in self._counters:
            return 0  # defaultdict(lambda: 0) behaviour of the reference
        return int(np.asarray(self._counters[variable])[0])

    def __contains__(self, variable):
        return variable in self._counters

    def __len__(self):
        return len(self._counters)


_table_cache = {}  # tuple of device pointers -> ctypes pointer table (buffers are recycled:
                   # a training loop cycles through a handful of distinct tuples)


def _cached_table(ptrs):
    tab = _table_cache.get(ptrs)
    if tab is None:
        if len(_table_cache) > 256:
            _table_cache.clear()
        tab = _table_cache[ptrs] = _ptr_table(ptrs)
    return tab


def _settle_before_inplace_update(gradients) -> None:
    """Parameters are about to be overwritten IN PLACE: everything that was deferred and would
    read them later must run now, or it would see the updated values (the reference evaluates
    every gradient inside get_gradients, sp.py:77-87)."""
    materialise = getattr(gradients, "_materialise", None)
    if materialise is not None:
        materialise()
    A.flush_pending_parameter_readers()


class Adam:
    """Adam optimization for SmallPebble variables (sp.py:552-583; Kingma & Ba, Algorithm 1).

    Per-variable step count / first / second moment, as in the reference.  One fused kernel
    updates every variable.  Like the reference (sp.py:583) the update is written to a NEW
    array that `variable.array` is rebound to: whatever still holds the old array -- a
    gradient closure that has not been evaluated yet (`Gradients` defers the gradient of
    non-learnable leaves until somebody reads it), a deferred forward op -- keeps seeing the
    weights the step was computed with.  Only while a CUDA graph is being recorded
    (`sp.capture`), where buffers must be static, the update is in place (pending work that
    reads the parameters is settled first)."""

    def __init__(self, alpha: float = 0.001, beta1: float = 0.9, beta2: float = 0.999,
                 eps: float = 1e-8) -> None:
        self.alpha = alpha
        self.beta1 = beta1
        self.beta2 = beta2
        self.eps = eps
        self.m = {}
        self.v = {}
        self._step = {}  # Variable -> device int32[1]: completed steps (read by the kernel)
        self.t = _StepCounts(self._step)
        self._static = None  # (variable ids, moment / step-count / size tables, variables)

    def training_step(self, variables, gradients) -> None:
        variables = list(variables)
        if not variables:
            return
        rec = getattr(gradients, "_steady", None)
        if rec is not None and _steady.optimiser_step(self, variables, gradients, rec):
            return  # replayed (or recorded and replayed) from the steady-state graphs
        self._update(variables, gradients, None)

    def _ensure_state(self, variables) -> None:
        "Moments and step counters of variables seen for the first time (on the device)."
        for v in variables:
            if v not in self._step:
                self.m[v] = DeviceArray.zeros(v.array.shape)
                self.v[v] = DeviceArray.zeros(v.array.shape)
                self._step[v] = DeviceArray.from_host(np.zeros(1, np.int32), dtype=np.int32)
                self._step[v].ptr  # uploaded now: never inside a graph capture

    def _hyper(self):
        return (self.alpha, self.beta1, self.beta2, self.eps)

    def _update(self, variables, gradients, outs) -> None:
        """The fused update.  `outs`: static arrays to write the new parameters into (steady.py,
        while the update is being recorded); None: fresh arrays (eager) or in place (sp.capture)."""
        grads = []
        for v in variables:
            g = gradients[v]  # KeyError when missing, like the reference
            if type(g) is not DeviceArray or g.shape != v._array.shape:
                g = _device_gradient(v, gradients)
            grads.append(g)
        peer = None
        if (_distributed.is_initialized() and _distributed.world_size() > 1
                and not getattr(gradients, "_reduced", False)):
            # small buckets: the sum over the ranks happens inside the update kernel, over
            # NVLink peer memory; otherwise NCCL all-reduces the bucket first
            peer = _distributed.peer_plan(grads)
            if peer is None:
                grads = _distributed.allreduce_gradients(grads, variables)
        n = len(variables)
        # state creation and the pointer tables of moments / step counters are keyed by the
        # identity of the variables: they never change once made
        ids = tuple([id(v) for v in variables])
        cached = self._static
        if cached is None or cached[0] != ids:
            self._ensure_state(variables)
            cached = self._static = (
                ids,
                _ptr_table([self.m[v].ptr for v in variables]),
                _ptr_table([self.v[v].ptr for v in variables]),
                _abi.dims([v.array.size for v in variables]),
                _ptr_table([self._step[v].ptr for v in variables]),
                tuple(variables),
            )
        _, m_tab, v_tab, numel, step_tab, _ = cached
        olds = [v._array for v in variables]
        p_tab = _cached_table(tuple([a.ptr for a in olds]))
        g_tab = _cached_table(tuple([g.ptr for g in grads]))
        if A.capturing() and outs is None:
            # static buffers: update in place.  (Nothing is settled first: every buffer of a
            # captured step is overwritten by the next replay anyway, so values that were not
            # evaluated inside the captured function cannot be read afterwards -- capture.py.)
            if peer is not None:
                F.sp_adam_multi_peer(peer[0], n, p_tab, p_tab, g_tab, peer[1], m_tab, v_tab, numel,
                                     step_tab, None, None, None, None, self.alpha, self.beta1,
                                     self.beta2, self.eps, A.stream())
            else:
                F.sp_adam_multi_out(n, p_tab, p_tab, g_tab, m_tab, v_tab, numel, step_tab,
                                    self.alpha, self.beta1, self.beta2, self.eps, A.stream())
            A.note_inplace_update()
            return
        if outs is None:
            news = [DeviceArray.empty(a.shape, A.FLOAT, a.size) for a in olds]
        else:
            news = outs
        o_tab = _cached_table(tuple([a._ptr for a in news]))
        # The new arrays get the TF32 shadows the contractions asked of the old ones (rounded
        # copy for the data gradients, split copy for the compensated forward pass), written
        # by the update kernel itself; a tensor with more than one split layout gets the rest
        # from one extra launch.
        refresh = outs is not None
        shadow_rows, rest_jobs = None, []
        for i, (old, new) in enumerate(zip(olds, news)):
            want_round, kinds = A.tf32_shadow_requests(old)
            if not (want_round or kinds) or not new.size:
                continue
            rounded, split, rest = A.tf32_shadow_targets(new, want_round, kinds, refresh)
            if rounded is not None or split is not None:
                if shadow_rows is None:
                    shadow_rows = [[None] * n, [None] * n, [1] * n, [0] * n]
                if rounded is not None:
                    shadow_rows[0][i] = rounded._ptr
                if split is not None:
                    shadow_rows[1][i] = split[0]._ptr
                    shadow_rows[3][i], shadow_rows[2][i] = split[1], split[2]
            if rest:
                rest_jobs.append((new, False, rest))
        if shadow_rows is None:
            if peer is not None:
                F.sp_adam_multi_peer(peer[0], n, p_tab, o_tab, g_tab, peer[1], m_tab, v_tab, numel,
                                     step_tab, None, None, None, None, self.alpha, self.beta1,
                                     self.beta2, self.eps, A.stream())
            else:
                F.sp_adam_multi_out(n, p_tab, o_tab, g_tab, m_tab, v_tab, numel, step_tab,
                                    self.alpha, self.beta1, self.beta2, self.eps, A.stream())
        else:
            key = (tuple(shadow_rows[0]), tuple(shadow_rows[1]), tuple(shadow_rows[2]),
                   tuple(shadow_rows[3]))
            tabs = _shadow_tables.get(key)
            if tabs is None:
                if len(_shadow_tables) > 64:
                    _shadow_tables.clear()
                tabs = _shadow_tables[key] = (
                    _ptr_table(key[0]), _ptr_table(key[1]), _abi.dims(key[2]),
                    (C.c_int32 * n)(*key[3]))
            if peer is not None:
                F.sp_adam_multi_peer(peer[0], n, p_tab, o_tab, g_tab, peer[1], m_tab, v_tab, numel,
                                     step_tab, tabs[0], tabs[1], tabs[2], tabs[3], self.alpha,
                                     self.beta1, self.beta2, self.eps, A.stream())
            else:
                F.sp_adam_multi_shadow(n, p_tab, o_tab, g_tab, m_tab, v_tab, numel, step_tab,
                                       tabs[0], tabs[1], tabs[2], tabs[3], self.alpha, self.beta1,
                                       self.beta2, self.eps, A.stream())
        for v, new in zip(variables, news):
            v._array = new  # rebind (sp.py:583)
        if rest_jobs:
            A.tf32_shadows_multi(rest_jobs, refresh=refresh)


_shadow_tables = {}  # (shadow pointers, layouts) of an update -> ctypes tables


def sgd_step(variables, gradients, learning_rate: float = 0.001) -> None:
    "A single step of gradient descent. Modifies each variable.array directly (sp.py:645-653)."
    from . import distributed

    variables = list(variables)
    if not variables:
        return
    grads = [_device_gradient(v, gradients) for v in variables]
    if (distributed.is_initialized() and distributed.world_size() > 1
            and not getattr(gradients, "_reduced", False)):
        grads = distributed.allreduce_gradients(grads, variables)
    # in place, like the reference: settle deferred work that reads the parameters first
    _settle_before_inplace_update(gradients)
    F.sp_sgd_multi(
        len(variables),
        _ptr_table([v.array.ptr for v in variables]),
        _ptr_table([g.ptr for g in grads]),
        _abi.dims([v.array.size for v in variables]),
        learning_rate, A.stream(),
    )
    A.note_inplace_update()


# ---------------- NEURAL NETWORK HELPERS (sp.py:586-642)


def batch(X: np.ndarray, y: np.ndarray, size: int, seed=None):
    """Yield sub-batches of X,y, randomly selecting with replacement (sp.py:586-593).

    Same draws from the global NumPy RNG and the same values as the reference.  With a GPU
    present a floating-point batch of 64 KB or more is gathered straight into a fresh
    page-locked array (the input pipeline): `sp.Variable(xbatch)` then uploads
    it by DMA without a staging pass and narrows float64 on the device."""
    assert (y.ndim == 1) and (X.shape[0] == y.size), "unexpected dimensions."
    if seed:
        np.random.seed(seed)
    X = np.asarray(X)
    row_bytes = (X.size // max(X.shape[0], 1)) * X.dtype.itemsize
    pin = (X.dtype in (np.float32, np.float64) and X.ndim >= 1
           and row_bytes * size >= A._PIN_THRESHOLD and X.flags.c_contiguous)
    while True:
        idx = np.random.randint(0, y.size, size)
        xb = A.pinned_empty((size,) + X.shape[1:], X.dtype) if pin else None
        if xb is not None:
            np.take(X, idx, axis=0, out=xb)
        else:
            xb = X[idx, ...]
        yield xb, y[idx]


def convlayer(height: int, width: int, depth: int, n_kernels: int, padding: str = "VALID",
              strides=(1, 1)):
    "Create a convolutional neural network layer (sp.py:596-609)."
    sigma = np.sqrt(6 / (height * width * depth + height * width * n_kernels))
    kernels_init = sigma * (np.random.random([height, width, depth, n_kernels]) - 0.5)
    kernels = learnable(Variable(kernels_init))
    func = lambda images, kernels: core.conv2d(images, kernels, padding, strides)  # noqa: E731
    return lambda images: Lazy(func)(images, kernels)


def he_init(insize: int, outsize: int) -> np.ndarray:
    "He weight initialization (sp.py:624-627)."
    sigma = np.sqrt(4 / (insize + outsize))
    return np.random.random([insize, outsize]) * sigma - sigma / 2


def linearlayer(insize: int, outsize: int):
    "Create a linear fully connected neural network layer (sp.py:630-635)."
    weights = learnable(Variable(he_init(insize, outsize)))
    bias = learnable(Variable(np.ones([outsize], np.float32)))
    func = lambda a, weights, bias: core.matmul(a, weights) + bias  # noqa: E731
    return lambda a: Lazy(func)(a, weights, bias)


def onehot(y: np.ndarray, n_classes: int) -> np.ndarray:
    "Onehot encode vector y with classes 0 to n_classes-1 (sp.py:638-642)."
    result = np.zeros([len(y), n_classes])
    result[np.arange(len(y)), y] = 1
    return result
