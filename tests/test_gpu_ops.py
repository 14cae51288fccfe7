"""GPU parity tests proper: every hot-path op, through the public API (which calls the C
ABI), against the oracle on the same seeded inputs AND against the golden vectors recorded
from the unmodified reference.  Integer results (argmax indices) bit-exact; floating point
within BASELINE.json's relative tolerance of the float64 values (1e-5 fp32, 2e-3 TF32)."""

import numpy as np
import pytest

import cases as C
import smallpebble_b200 as sp
from conftest import FWD_RTOL, RTOL, assert_close
from oracle import numpy_path as orc
from smallpebble_b200 import _abi, core

pytestmark = pytest.mark.gpu


def grads_of(fn, arrays, gy):
    vs = [sp.Variable(a) for a in arrays]
    y = fn(*vs)
    g = sp.get_gradients(sp.mul(y, sp.Variable(gy)))
    return y, [g[v] for v in vs]


# sp_debug_set_umma_policy: bits 0-1 kernel family, bits 5-6 the 2-SM TMA conv kernels
# (32 = whenever the geometry is eligible, 64 = never)
# bits 7-8: the halo-reuse variant of the 2-SM kernels (128 = whenever eligible)
@pytest.fixture(params=[0, 1, 2, 32, 1 + 64, 32 + 128],
                ids=["policy", "force_umma", "force_simt", "force_tma2", "umma_no_tma2",
                     "force_halo"])
def umma_policy(request):
    core.debug_set_kernel_policy(request.param)
    yield request.param
    core.debug_set_kernel_policy(0)


@pytest.mark.parametrize("case", C.CONV_CASES, ids=[c[0] for c in C.CONV_CASES])
def test_conv2d(case, precision, umma_policy, golden):
    if precision == "fp32" and umma_policy not in (0, 2):
        pytest.skip("tcgen05 kernels are only selected in TF32 mode")
    if precision == "tf32" and umma_policy == 32 + 128:
        pytest.skip("the halo kernel has no split-operand mode: compensated forward passes "
                    "take the im2col 2-SM kernel (force_tma2 covers it)")
    name, xs, ws, strides, pad = case
    x, w, gy = C.conv_inputs(case)
    tol = RTOL[precision]
    y, (dx, dw) = grads_of(lambda a, b: sp.conv2d(a, b, pad, strides), [x, w], gy)
    y0, dx0, dw0 = orc.conv2d_grads(x, w, gy, pad, strides)
    assert y.shape == y0.shape
    assert_close(y.array, y0, FWD_RTOL[precision], "y vs oracle")
    assert_close(dx, dx0, tol, "dX vs oracle")
    assert_close(dw, dw0, tol, "dW vs oracle")
    assert_close(y.array, golden[f"conv/{name}/y"], FWD_RTOL[precision], "y vs golden")
    assert_close(dx, golden[f"conv/{name}/dx"], tol, "dX vs golden")
    assert_close(dw, golden[f"conv/{name}/dw"], tol, "dW vs golden")


@pytest.mark.parametrize("case", C.POOL_CASES, ids=[c[0] for c in C.POOL_CASES])
def test_maxpool2d_bit_exact_indices(case, golden):
    name, xs, (kh, kw), strides, pad = case
    x, gy = C.pool_inputs(case)
    x32 = x.astype(np.float32)  # the device sees fp32 inputs: argmax is compared on those
    y, (dx,) = grads_of(lambda a: sp.maxpool2d(a, kh, kw, pad, list(strides)), [x], gy)
    y0, idx0 = orc.maxpool2d_forward(x32, kh, kw, pad, strides)
    _, dx0 = orc.maxpool2d_grads(x32, kh, kw, gy.astype(np.float32), pad, strides)
    idx = np.asarray(y._argmax)
    assert idx.dtype == np.uint8
    assert np.array_equal(idx, idx0)  # bit-exact index map
    assert np.array_equal(np.asarray(y.array), y0.astype(np.float32))  # a max is exact
    assert_close(dx, dx0, 1e-6, "dX vs oracle")
    # the float64 golden argmax agrees wherever fp32 rounding created no new ties
    assert np.mean(idx == golden[f"pool/{name}/idx"]) > 0.999
    assert_close(y.array, golden[f"pool/{name}/y"], 1e-6, "y vs golden")
    assert_close(dx, golden[f"pool/{name}/dx"], 1e-5, "dX vs golden")


def test_maxpool2d_padding_ties_nan():
    x = -np.ones((1, 3, 3, 4))
    v = sp.Variable(x)
    y = sp.maxpool2d(v, 2, 2, "SAME", [2, 2])
    out = np.asarray(y.array)
    assert np.all(out[0, 1, 1] == 0.0) and np.all(out[0, 0, 0] == -1.0)  # zero pad wins
    assert np.all(np.asarray(y._argmax)[0, 0, 0] == 0)  # all-equal window -> index 0
    dx = np.asarray(sp.get_gradients(y)[v])
    assert dx.sum() == 4.0  # windows won by the pad drop their gradient
    x = np.zeros((1, 2, 2, 3))
    x[0, 1, 0, :] = np.nan
    y = sp.maxpool2d(sp.Variable(x), 2, 2, "VALID", [2, 2])
    assert np.all(np.isnan(np.asarray(y.array))) and np.all(np.asarray(y._argmax) == 2)


@pytest.mark.parametrize("case", C.MATMUL_CASES, ids=[c[0] for c in C.MATMUL_CASES])
def test_matmul(case, precision, umma_policy, golden):
    if precision == "fp32" and umma_policy != 0 and umma_policy != 2:
        pytest.skip("tcgen05 kernels are only selected in TF32 mode")
    if umma_policy >= 32:
        pytest.skip("the 2-SM TMA kernels are conv-only")
    a, b, gy = C.matmul_inputs(case)
    tol = RTOL[precision]
    y, (da, db) = grads_of(sp.matmul, [a, b], gy)
    y0, da0, db0 = orc.matmul_grads(a, b, gy)
    assert_close(y.array, y0, FWD_RTOL[precision], "y")
    assert_close(da, da0, tol, "dA")
    assert_close(db, db0, tol, "dB")
    assert_close(y.array, golden[f"matmul/{case[0]}/y"], FWD_RTOL[precision], "y vs golden")
    assert_close(da, golden[f"matmul/{case[0]}/da"], tol, "dA vs golden")
    assert_close(db, golden[f"matmul/{case[0]}/db"], tol, "dB vs golden")


@pytest.mark.parametrize("case", C.SOFTMAX_CASES, ids=[c[0] for c in C.SOFTMAX_CASES])
def test_softmax(case, golden):
    name, shape, axis = case
    x, gy = C.softmax_inputs(case)
    y, (dx,) = grads_of(lambda a: sp.softmax(a, axis), [x], gy)
    assert_close(y.array, golden[f"softmax/{name}/y"], 1e-5, "y")
    assert_close(dx, golden[f"softmax/{name}/dx"], 1e-5, "dx")


@pytest.mark.parametrize("case", C.MAXAX_CASES, ids=[c[0] for c in C.MAXAX_CASES])
def test_maxax(case, golden):
    """Values AND gradient layout of the reference for axis != last (sp.py:257-279: the result is
    reshaped without swapping back; the gradient multiplies the upstream gradient in its natural
    layout by the swapped-back one-hot) -- ADVICE r01."""
    name, shape, axis = case
    x, gy = C.maxax_inputs(case)
    y, (dx,) = grads_of(lambda a: sp.maxax(a, axis), [x], gy)
    assert y.shape == golden[f"maxax/{name}/y"].shape
    assert_close(y.array, golden[f"maxax/{name}/y"], 1e-6, "y")
    assert_close(dx, golden[f"maxax/{name}/dx"], 1e-6, "dx")


@pytest.mark.parametrize("fused", [True, False])
def test_softmax_cross_entropy(fused, golden):
    sp.set_fusion(fused)
    try:
        logits, labels = C.xent_inputs()
        v = sp.Variable(logits)
        p = sp.softmax(v)
        loss = sp.cross_entropy(p, labels)
        assert loss.array.shape == ()
        g = sp.get_gradients(loss)
        assert getattr(loss, "_fused_softmax_xent", False) == fused
        assert_close(p.array, golden["xent/probs"], 1e-5, "probs")
        assert_close(loss.array, golden["xent/loss"], 1e-5, "loss")
        assert_close(g[v], golden["xent/dlogits"], 1e-5, "dlogits")
        if not fused:
            assert p in g  # the probability node's own gradient is reported
    finally:
        sp.set_fusion(True)


def test_cross_entropy_known_answer():
    p = sp.Variable(np.full((2, 3), 1 / 3))
    loss = sp.cross_entropy(p, np.array([0, 2]))
    assert abs(float(loss.array) - 2 * np.log(3)) < 1e-5
    with pytest.raises(IndexError):
        sp.cross_entropy(p, np.array([0, 3]))


def test_leaky_relu(golden):
    x, gy = C.leaky_inputs()
    for alpha in (0.02, 0.1):
        y, (dx,) = grads_of(lambda a: sp.leaky_relu(a, alpha), [x], gy)
        assert_close(y.array, golden[f"leaky/{alpha}/y"], 1e-6, "y")
        assert_close(dx, golden[f"leaky/{alpha}/dx"], 1e-6, "dx")
    v = sp.Variable(np.array([-1.0, 0.0, 2.0]))
    y = sp.leaky_relu(v)
    assert np.allclose(np.asarray(y.array), [-0.02, 0.0, 2.0])
    assert np.allclose(np.asarray(sp.get_gradients(y)[v]), [0.02, 0.02, 1.0])


def test_bias_add_and_unbroadcast_sum(golden):
    a, bias, gy = C.bias_inputs()
    y, (da, db) = grads_of(sp.add, [a, bias], gy)
    assert_close(y.array, golden["bias/y"], 1e-6, "y")
    assert_close(da, golden["bias/da"], 1e-6, "da")
    assert_close(db, golden["bias/db"], 1e-5, "dbias")
    assert np.asarray(db).shape == (100,)


def test_adam_and_sgd_steps(golden):
    p0, grads = C.adam_inputs()
    v = sp.Variable(p0.copy())
    adam = sp.Adam()
    for i, g in enumerate(grads):
        adam.training_step([v], {v: g})
        assert_close(v.array, golden[f"adam/p{i + 1}"], 1e-5, f"adam step {i + 1}")
    v = sp.Variable(np.array([1.0, -2.0, 3.0]))
    adam = sp.Adam()
    g = np.array([0.1, -0.2, 0.3])
    adam.training_step([v], {v: g})
    assert np.allclose(np.asarray(v.array), [0.999, -1.999, 2.999], atol=1e-6)
    adam.training_step([v], {v: g})
    assert np.allclose(np.asarray(v.array), [0.998, -1.998, 2.998], atol=1e-6)
    with pytest.raises(KeyError):
        adam.training_step([sp.Variable(np.ones(2))], {})
    v = sp.Variable(p0.copy())
    sp.sgd_step([v], {v: grads[0]}, learning_rate=0.01)
    assert_close(v.array, golden["sgd/p1"], 1e-6, "sgd")


def test_adam_many_tensors_odd_sizes():
    rng = np.random.default_rng(5)
    shapes = [(3,), (17, 5), (4097,), (1,), (64, 33)] * 6  # 30 tensors: two kernel batches
    vs = [sp.Variable(rng.random(s) - 0.5) for s in shapes]
    nodes = [orc.Node(np.asarray(v.array, dtype=np.float64)) for v in vs]
    adam, opt = sp.Adam(alpha=0.01), orc.AdamState(alpha=0.01)
    for _ in range(3):
        gs = [rng.random(s) - 0.5 for s in shapes]
        adam.training_step(vs, dict(zip(vs, gs)))
        opt.step(nodes, dict(zip(nodes, gs)))
    for v, n in zip(vs, nodes):
        assert_close(v.array, n.value, 2e-5, "multi-tensor adam")


@pytest.mark.parametrize("pool", [((2, 2), (2, 2), "SAME"), ((3, 3), (3, 3), "SAME"),
                                  ((2, 2), (1, 1), "VALID"), ((3, 2), (2, 3), "SAME")],
                         ids=["k2s2", "k3s3", "k2s1_overlap", "k3x2_s2x3"])
@pytest.mark.parametrize("alpha", [0.02, 0.0, -0.5])
def test_leaky_relu_maxpool_fusion_is_bit_identical(pool, alpha):
    """maxpool2d(leaky_relu(op_result)) runs as ONE deferred-launch kernel forward and ONE
    backward (core.leaky_relu docstring).  Values, argmax indices and gradients must equal the
    unfused path bit for bit -- including alpha = 0 (ties between clamped values) and a
    negative slope (non-monotonic activation), and the skipped leaky output must still
    materialise correctly when somebody reads it."""
    (kh, kw), strides, pad = pool
    rng = np.random.default_rng(21)
    x = rng.random((3, 9, 11, 8)) - 0.5
    gy_seed = rng.random((3, 9, 11, 8)) + 0.5

    def run(fuse):
        sp.set_fusion(fuse)
        try:
            xv = sp.Variable(x)
            pre = sp.mul(xv, sp.Variable(np.full(x.shape, 1.5)))   # an op result feeds leaky
            act = sp.leaky_relu(pre, alpha)
            pooled = sp.maxpool2d(act, kh, kw, pad, list(strides))
            assert (act.array.pending_tag is not None) == fuse
            gy = rng.random(pooled.shape)
            g = sp.get_gradients(sp.mul(pooled, sp.Variable(np.ones(pooled.shape) * 0.75)))
            out = (np.asarray(pooled.array), np.asarray(pooled._argmax), np.asarray(g[xv]),
                   np.asarray(g[pre]), np.asarray(g[act]), np.asarray(act.array))
            del gy
            return out
        finally:
            sp.set_fusion(True)

    fused, plain = run(True), run(False)
    for a, b, what in zip(fused, plain, ("pooled", "argmax", "dX", "d(pre)", "d(act)", "act")):
        assert np.array_equal(a, b, equal_nan=True), f"{what} differs between fused and unfused"
    # and against the oracle
    x32 = (x.astype(np.float32) * np.float32(1.5))
    act0 = np.where(x32 > 0, x32, x32 * np.float32(alpha))
    y0, idx0 = orc.maxpool2d_forward(act0, kh, kw, pad, strides)
    assert np.array_equal(fused[0], y0.astype(np.float32)) and np.array_equal(fused[1], idx0)
    del gy_seed


_VIEW_INDICES = [1, -1, slice(1, 4), slice(None, None, 2), slice(4, 0, -1), (1, 2), (slice(1, 3), 0),
                 (Ellipsis, 2), (None, 1), (slice(None), None, slice(2, 6, 3)),
                 (0, Ellipsis, slice(None, None, -2)), slice(3, 3), (2, 3, 4),
                 (np.array([0, 2, 2, 4]), slice(1, 3)), (np.arange(5), np.arange(5), 0)]


@pytest.mark.parametrize("index", _VIEW_INDICES, ids=[str(i)[:40] for i in _VIEW_INDICES])
def test_getitem_add_at_setat_match_numpy(index):
    """f3: basic indices run as strided copies straight from the view geometry, advanced ones
    through an index map; both must be NumPy's `a[index]`, `np.add.at` and `a[index] = b`
    (sp.py:111-121, 210-221, 340-354), forward and gradient."""
    rng = np.random.default_rng(3)
    x = rng.random((5, 6, 7)).astype(np.float32)
    picked = x[index]
    gy = rng.random(picked.shape).astype(np.float32)
    y, (dx,) = grads_of(lambda a: sp.getitem(a, index), [x], gy)
    assert y.shape == picked.shape
    assert np.array_equal(np.asarray(y.array), picked)
    want = np.zeros_like(x)
    np.add.at(want, index, gy)
    assert_close(dx, want, 1e-6, "getitem gradient")
    b = rng.random(picked.shape).astype(np.float32)
    gfull = rng.random(x.shape).astype(np.float32)
    y, (da, db) = grads_of(lambda a, c: sp.add_at(a, index, c), [x, b], gfull)
    want = x.copy()
    np.add.at(want, index, b)
    assert_close(y.array, want, 1e-6, "add_at")
    assert np.array_equal(np.asarray(da), gfull)
    assert np.array_equal(np.asarray(db), gfull[index])
    y, (da, db) = grads_of(lambda a, c: sp.setat(a, index, c), [x, b], gfull)
    want = x.copy()
    want[index] = b
    if not isinstance(index, tuple) or not any(isinstance(i, np.ndarray) for i in index):
        assert np.array_equal(np.asarray(y.array), want)  # (repeated advanced indices: any writer may win)
    keep = gfull.copy()
    keep[index] = 0
    assert np.array_equal(np.asarray(da), keep)
    assert np.array_equal(np.asarray(db), gfull[index])


@pytest.mark.parametrize("shape,window,strides", [((9, 10), (3, 4), (2, 3)),
                                                  ((2, 7, 8, 3), (1, 3, 2, 1), (1, 2, 3, 1)),
                                                  ((6,), (6,), (1,))])
def test_strided_sliding_view_matches_numpy(shape, window, strides):
    rng = np.random.default_rng(4)
    x = rng.random(shape).astype(np.float32)
    ref = core.np_strided_sliding_view(x, window, strides)
    gy = rng.random(ref.shape).astype(np.float32)
    y, (dx,) = grads_of(lambda a: sp.strided_sliding_view(a, window, strides), [x], gy)
    assert np.array_equal(np.asarray(y.array), ref)
    want = np.zeros(x.shape, np.float64)
    idx = core.np_strided_sliding_view(np.arange(x.size).reshape(shape), window, strides)
    np.add.at(want.reshape(-1), idx.reshape(-1), gy.reshape(-1).astype(np.float64))
    assert_close(dx, want, 1e-6, "overlapping windows accumulate")
