"""CPU: host-side logic of the product package (no kernel runs): integer helpers bit-exact
with the golden vectors, lazy graph / learnables behaviour (the reference's own tests
344-348, 401-416), error behaviour, lazy upload, and loud failure without a GPU."""

import numpy as np
import pytest
import torch

import smallpebble_b200 as sp
from smallpebble_b200 import _abi
from smallpebble_b200 import distributed as dist

NO_GPU = not torch.cuda.is_available()


def test_pad_amounts_bit_exact(golden):
    pads = [sp.pad_amounts(L, s, k) for L in range(1, 40) for s in (1, 2, 3, 4, 7, 10)
            for k in (1, 2, 3, 5, 13, 22)]
    assert np.array_equal(np.array(pads, np.int64), golden["int/pad_amounts"])


def test_patches_index_bit_exact(golden):
    for args in [(5, 6, 2, 3, 2, 1), (32, 32, 3, 3, 1, 1), (13, 17, 5, 2, 3, 4)]:
        (r, c), oh, ow, n = sp.patches_index(*args)
        key = "int/patches_index/" + "_".join(map(str, args))
        assert r.dtype == np.int64 and c.dtype == np.int64
        assert np.array_equal(r, golden[key + "/rows"])
        assert np.array_equal(c, golden[key + "/cols"])
        assert [oh, ow, n] == list(golden[key + "/meta"])


def test_broadcastinfo():
    assert sp.broadcastinfo((4, 2, 3), (3, 4, 1, 2, 3)) == ((0, 1), (2,))
    assert sp.broadcastinfo((), (3, 1)) == ((0, 1), ())
    with pytest.raises(ValueError, match="could not broadcast"):
        sp.broadcastinfo((2, 3), (4, 3))


def test_np_strided_sliding_view_matches_numpy():
    x = np.arange(5 * 7).reshape(5, 7)
    v = sp.np_strided_sliding_view(x, (2, 3), (2, 1))
    ref = np.lib.stride_tricks.sliding_window_view(x, (2, 3))[::2, ::1]
    assert np.array_equal(v, ref)
    with pytest.raises(ValueError):
        sp.np_strided_sliding_view(x, (2,), (1, 1))
    with pytest.raises(ValueError):
        sp.np_strided_sliding_view(x, (6, 1), (1, 1))


def test_operation_and_placeholder_basic():  # ref tests:344-348
    x_in = sp.Placeholder()
    x = sp.Lazy(lambda x: x**2)(x_in)
    x_in.assign_value(3)
    assert x.run() == 9


def test_lazy_errors_and_no_memoisation():
    node = sp.Lazy(lambda a: a)
    with pytest.raises(sp.AssignmentError):
        node.run()
    node(1)
    with pytest.raises(sp.AssignmentError):
        node(2)
    calls = []
    shared = sp.Lazy(lambda a: calls.append(a) or a)(5)
    top = sp.Lazy(lambda a, b: a + b)(shared, shared)
    assert top.run() == 10 and len(calls) == 2  # evaluated once per parent


def test_learnable_and_get_learnables():  # ref tests:401-416
    param_1 = sp.learnable(sp.Variable(np.array(5)))
    param_2 = sp.learnable(sp.Variable(np.array(10)))
    a, b = sp.Placeholder(), sp.Placeholder()
    y = sp.Lazy(sp.matmul)(a, b)
    y = sp.Lazy(sp.mul)(y, param_1)
    y = sp.Lazy(sp.add)(y, param_2)
    assert sp.get_learnables(y) == [param_1, param_2]


def test_initialisers_match_reference_rng_recipe():
    """convlayer / linearlayer draw from the global NumPy RNG exactly like sp.py:605-606,
    624-633, so the oracle's models start from bit-identical float64 weights."""
    from oracle import numpy_path as orc

    np.random.seed(0)
    x_in = sp.Placeholder()
    h = sp.convlayer(3, 3, 3, 32)(x_in)
    h = sp.linearlayer(20, 10)(h)
    params = sp.get_learnables(h)
    np.random.seed(0)
    k = orc.conv_kernel_init(3, 3, 3, 32)
    w, b = orc.linear_init(20, 10)
    got = [np.asarray(p.array) for p in params]
    assert [g.shape for g in got] == [k.shape, w.shape, b.shape]
    for g, e in zip(got, (k, w, b)):
        assert np.array_equal(g, e.astype(np.float32))  # same draw, cast once to fp32


def test_variable_holds_lazy_device_array():
    a = np.arange(6, dtype=np.float64).reshape(2, 3)
    v = sp.Variable(a)
    assert v.shape == (2, 3) and v.ndim == 2 and v.dtype == np.float32
    assert not v.array.on_device
    a[0, 0] = 99.0  # the Variable took a snapshot
    assert np.asarray(v.array)[0, 0] == 0.0
    assert v.array.reshape(3, 2).shape == (3, 2)
    with pytest.raises(ValueError):
        v.array.reshape(4, 2)


def test_conv_and_pool_argument_errors_before_any_kernel():
    x = sp.Variable(np.zeros((1, 4, 4, 2)))
    w = sp.Variable(np.zeros((3, 3, 2, 5)))
    with pytest.raises(ValueError, match="SAME"):
        sp.conv2d(x, w, padding="FULL")
    with pytest.raises(ValueError, match="SAME"):
        sp.maxpool2d(x, 2, 2, padding="nope")
    with pytest.raises(ValueError):
        sp.conv2d(x, sp.Variable(np.zeros((3, 3, 3, 5))))
    with pytest.raises(ValueError):
        sp.conv2d(x, sp.Variable(np.zeros((5, 5, 2, 5))), padding="VALID")


def test_batch_sampler_follows_global_rng():
    X, y = np.arange(40).reshape(10, 4), np.arange(10)
    xb, yb = next(sp.batch(X, y, 6, seed=3))
    np.random.seed(3)
    idx = np.random.randint(0, 10, 6)
    assert np.array_equal(xb, X[idx]) and np.array_equal(yb, y[idx])
    with pytest.raises(AssertionError):
        next(sp.batch(X, y.reshape(5, 2), 2))


def test_onehot():
    assert np.array_equal(sp.onehot(np.array([2, 0]), 3), [[0, 0, 1], [1, 0, 0]])


def test_shard_bounds_partition_every_row_once():
    for n in (0, 1, 7, 128, 1024, 1031):
        for world in (1, 2, 3, 8):
            cuts = [dist.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_bucket_layout_is_16_byte_aligned():
    sizes = [864, 36864, 147456, 11520, 10, 3]
    lay = dist.bucket_layout(sizes)
    assert lay[0] == 0 and all(o % 4 == 0 for o in lay)
    assert all(lay[i + 1] - lay[i] >= sizes[i] for i in range(len(sizes)))


@pytest.mark.skipif(not NO_GPU, reason="only meaningful on a machine without CUDA")
def test_compute_fails_loudly_without_gpu():
    v = sp.Variable(np.ones((2, 2)))
    with pytest.raises(_abi.SmallPebbleCudaError, match="no CPU fallback"):
        sp.add(v, v)
    with pytest.raises(_abi.SmallPebbleCudaError):
        sp.get_gradients(v)


# ---- host-side machinery added for the launch-bound README loop (no kernel runs) -------------


def _random_tape(rng, n_leaves=4, n_ops=14):
    """A random DAG of Variables with dummy gradient closures (never called here)."""
    from smallpebble_b200 import core

    nodes = [core.Variable(np.zeros(1)) for _ in range(n_leaves)]
    for _ in range(n_ops):
        k = int(rng.integers(1, 4))
        children = [nodes[int(i)] for i in rng.integers(0, len(nodes), k)]
        nodes.append(core.Variable(np.zeros(1), [(c, None) for c in children]))
    return nodes


def test_tape_order_puts_every_parent_before_its_children():
    """get_gradients sweeps the tape in descending construction order (core._tape_order);
    on random DAGs that order must be a reverse topological order covering exactly the op
    nodes the iterator-based DFS (the fallback) reaches."""
    from smallpebble_b200 import core

    rng = np.random.default_rng(5)
    for _ in range(50):
        nodes = _random_tape(rng)
        root = nodes[-1]
        order = core._tape_order(root)
        assert {id(n) for n in order} == {id(n) for n in core._tape_order_dfs(root)}
        position = {id(n): i for i, n in enumerate(order)}
        for n in order:
            for child, _ in n.local_gradients:
                if child.local_gradients:
                    assert position[id(n)] < position[id(child)]
    # a Variable that bypassed __init__ (no serial number) falls back to the DFS
    odd = core.Variable.__new__(core.Variable)
    odd._array, odd.local_gradients = nodes[0].array, [(nodes[0], None)]
    top = core.Variable(np.zeros(1), [(odd, None)])
    assert [id(n) for n in core._tape_order(top)] == [id(top), id(odd)]


def test_lazy_run_decides_once_per_argument_tuple_and_follows_reassignment():
    calls = []
    a, b = sp.Placeholder(), sp.Placeholder()
    node = sp.Lazy(lambda x, y, k: calls.append((x, y, k)) or x + y + k)(a, b, 10)
    a.assign_value(1)
    b.assign_value(2)
    assert node.run() == 13
    a.assign_value(5)                       # placeholders are re-read on every run
    assert node.run() == 17
    inner = sp.Lazy(lambda x: x * 2)(a)
    outer = sp.Lazy(lambda x, y: x - y)(inner, b)
    assert outer.run() == 8
    with pytest.raises(sp.AssignmentError):
        sp.Lazy(lambda: 0).run()
    with pytest.raises(sp.AssignmentError):
        sp.Placeholder().run()


def test_shape_plans_are_cached_and_validated():
    from smallpebble_b200 import array as A
    from smallpebble_b200 import core

    assert core._matmul_plan((5, 4, 3), (3, 2))[:5] == (4, 3, 2, (5,), ())
    assert core._matmul_plan((2, 1, 4, 3), (7, 3, 6))[7:] == ((2, 7), 14)
    with pytest.raises(ValueError, match="not aligned"):
        core._matmul_plan((4, 3), (2, 5))
    with pytest.raises(ValueError, match="at least 2 dimensions"):
        core._matmul_plan((3,), (3, 2))
    assert A._reshape_plan((-1, 6), 24) == (4, 6)
    with pytest.raises(ValueError, match="cannot reshape"):
        A._reshape_plan((5, 5), 24)
    steps, final, any_axis = A._reduce_plan((4, 5, 6), (0, 1), False)
    assert steps == ((1, 20, 6),) and final == (6,) and any_axis
    steps, final, _ = A._reduce_plan((4, 5, 6), (0, 2), True)
    assert steps == ((20, 6, 1), (1, 4, 5)) and final == (1, 5, 1)
    geom, out_shape, n_out = core._pool_plan((2, 7, 7, 3), 2, 2, 2, 2, "SAME")
    assert out_shape == (2, 4, 4, 3) and n_out == 96 and geom[-2:] == (4, 4)


def test_block_pool_recycles_on_last_reference_only():
    """csrc/fastcall.c: a Block hands its buffer to the per-size free list when the LAST
    reference dies (views share the Block), pool_alloc pops it, bypass and trim work.  Any
    Python object stands in for the torch tensor here."""
    fast = _abi._load_fast()
    if fast is None or not hasattr(fast, "pool_alloc"):
        pytest.skip("_sp_fastcall not built")
    fast.pool_trim()
    assert fast.pool_alloc(4096) is None
    payload = object()
    blk = fast.block_new(payload, 0xABC000, 4096, 1)
    assert (blk.ptr, blk.nbytes, blk.tensor) == (0xABC000, 4096, payload)
    view = blk                                  # a second holder
    del blk
    assert fast.pool_stats() == (0, 0)          # still referenced: not recycled
    del view
    assert fast.pool_stats() == (1, 4096)
    again = fast.pool_alloc(4096)
    assert again.ptr == 0xABC000 and again.tensor is payload and fast.pool_stats() == (0, 0)
    assert fast.pool_alloc(8192) is None        # exact size match only
    fast.pool_bypass(1)
    del again
    assert fast.pool_stats() == (0, 0)          # bypassed (CUDA-graph capture): freed, not pooled
    fast.pool_bypass(-1)
    once = fast.block_new(payload, 0x1000, 64, 0)
    del once
    assert fast.pool_stats() == (0, 0)          # recycle flag off
    keep = fast.block_new(payload, 0x2000, 64, 1)
    del keep
    fast.pool_trim()
    assert fast.pool_stats() == (0, 0)


def test_data_parallel_bucket_split_follows_production_order():
    """distributed._plan_split: the early part is everything but the last-produced <= 10 % of
    the bucket (at least the very last gradient); offsets stay 16-byte aligned and cover the
    bucket exactly once."""
    class G:
        def __init__(self, size):
            self.size, self.shape = size, (size,)

    variables = [object() for _ in range(5)]          # learnables in get_learnables order
    sizes = [864, 36864, 147456, 11520, 10]           # conv1, conv2, conv3, fc w, fc b
    grads = [G(n) for n in sizes]
    dist._overlap["written"].clear()
    for seq, i in enumerate([2, 3, 4, 1, 0], start=1):
        dist._overlap["written"][variables[i]] = (seq, id(grads[i]))
    try:
        split = dist._plan_split({}, grads, variables)
    finally:
        dist._overlap["written"].clear()
    assert split["order"] == [2, 3, 4, 1, 0] and split["k"] == 4
    assert split["mark"] is variables[1]              # conv2's gradient completes the early part
    assert all(off % 4 == 0 for off in split["offsets"])
    spans = sorted((split["offsets"][i], split["offsets"][i] + (sizes[i] + 3) // 4 * 4)
                   for i in range(5))
    assert spans[0][0] == 0 and spans[-1][1] == split["total"]
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert split["early_count"] + split["late_count"] == split["total"]
    assert split["late_count"] == 864                 # conv1's filters alone
    assert dist._plan_split({}, grads[:1], variables[:1]) is None
