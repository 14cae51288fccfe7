"""GPU: model-level parity.  The README MNIST MLP and CIFAR CNN, written with the public API
exactly as the README writes them, trained for k Adam steps from the reference's own initial
weights (same NumPy seed) on the same synthetic batch: loss curve, first-step gradients,
final weights and predictions against the golden vectors of the unmodified reference, and
against the oracle run in the same process."""

import numpy as np
import pytest

import cases as C
import smallpebble_b200 as sp
from conftest import assert_close, rel_l2
from oracle import numpy_path as orc
from smallpebble_b200 import workloads

pytestmark = pytest.mark.gpu

# Model-level tolerances.  BASELINE.json: fp32 / TF32 forward, gradients and Adam steps within
# 1e-5 / 2e-3 of the float64 reference.  A model gradient is the end of a CHAIN of contractions
# with discrete routing in between (maxpool argmax, leaky_relu sign).  The default "tf32" mode
# meets 2e-3 at model level because its forward contractions are error-compensated (routing
# identical to float64) and its backward operands are rounded, not truncated
# (scripts/tf32_error_attribution.py predicts ~7e-4; csrc/tf32.cu).  "tf32x1" (every
# contraction one raw tcgen05 pass) is the fast mode with a documented looser model-level
# bound: ~1e-4 of the routing decisions flip and the gradients -- sums with heavy
# cancellation at initialisation -- move by a few percent (the same script predicts 3-5e-2).
LOSS_RTOL = {"fp32": 2e-5, "tf32": 2e-5, "tf32x1": 2e-3}
GRAD_RTOL = {"fp32": 5e-5, "tf32": 2e-3, "tf32x1": 8e-2}
GRAD_L2 = {"fp32": 2e-5, "tf32": 2e-3, "tf32x1": 5e-2}
WEIGHT_RTOL = {"fp32": 1e-4, "tf32": 2e-3, "tf32x1": 5e-3}
# weights after k Adam steps, relative L2 over the well-conditioned elements: Adam divides the
# gradient's size out, so what is left of a gradient error is amplified by ~|g|max/|g| on the
# smaller elements -- fp32 turns its 2e-5 of gradient error into 1e-3 here, tf32 its ~6e-4 into
# ~3e-3 (measured; the per-element bound is the 2*alpha*steps envelope asserted beside it)
WEIGHT_L2 = {"fp32": 1e-3, "tf32": 5e-3, "tf32x1": 5e-2}


def assert_grad_close(got, expected, precision, what):
    assert_close(got, expected, GRAD_RTOL[precision], what)
    err = rel_l2(got, expected)
    assert err <= GRAD_L2[precision], f"{what}: relative L2 error {err:.3e}"


@pytest.mark.parametrize("kind,batch,steps", C.MODEL_CASES)
def test_training_matches_reference(kind, batch, steps, precision, golden):
    x, y = C.synthetic_batch(kind, batch)
    wl = workloads.BUILDERS[kind](seed=0)
    adam = sp.Adam()
    losses = []
    for step in range(steps):
        loss_val, grads = wl.train_step(adam, x, y)
        assert not np.isnan(loss_val.array)  # the README loop's own check
        if step == 0:
            for i, p in enumerate(wl.learnables):
                got = np.asarray(grads[p]).ravel()[::7]
                assert_grad_close(got, golden[f"model/{kind}/grad0/{i}"], precision,
                                  f"first-step gradient of learnable {i}")
        losses.append(float(loss_val.array))
    assert_close(losses, golden[f"model/{kind}/losses"], LOSS_RTOL[precision], "loss curve")
    for i, p in enumerate(wl.learnables):
        # Adam moves a weight by <= alpha per step whatever the gradient's size, and by
        # alpha*sign(g) while |g| dominates eps: an element whose gradient is ~0 can differ by
        # 2*alpha*steps between two correct implementations.  So: every element inside that
        # envelope, the tensor as a whole close in relative L2.
        got = np.asarray(p.array, np.float64).ravel()[::7]
        want = golden[f"model/{kind}/w_final/{i}"]
        assert np.max(np.abs(got - want)) <= 2.05e-3 * steps
        # well-conditioned elements (reference gradient clearly non-zero: the sign Adam divides
        # out is not rounding noise) must agree to the weight tolerance
        g0 = np.abs(golden[f"model/{kind}/grad0/{i}"])
        solid = g0 > min(5 * GRAD_RTOL[precision], 0.2) * g0.max()
        assert solid.any()
        err = rel_l2(got[solid], want[solid])
        assert err <= WEIGHT_L2[precision], f"weights of learnable {i}: relative L2 {err:.3e}"
    pred = wl.predict(x)
    assert_close(pred.array, golden[f"model/{kind}/pred_final"], 20 * WEIGHT_RTOL[precision],
                 "predictions after training")
    assert np.array_equal(np.argmax(np.asarray(pred.array), 1).shape, (batch,))


def assert_first_adam_step_close(p_new, p0_new, g0, rtol, grad_rtol, alpha=1e-3):
    """First Adam step: w -= alpha * g / (|g| + eps') = alpha * sign(g) for all but tiny g.
    Elements whose reference gradient is ~0 are ill-conditioned (their sign is rounding
    noise): they are only required to stay inside the 2*alpha envelope; every
    well-conditioned element must match to the weight tolerance."""
    p_new, p0_new, g0 = (np.asarray(t, np.float64) for t in (p_new, p0_new, g0))
    diff = np.abs(p_new - p0_new)
    assert diff.max() <= 2.05 * alpha
    solid = np.abs(g0) > min(5 * grad_rtol, 0.2) * np.abs(g0).max()
    assert solid.any()
    scale = max(np.abs(p0_new).max(), 1e-30)
    assert (diff[solid] / scale).max() <= rtol, (diff[solid] / scale).max()


def test_cnn_batch128_one_step_vs_oracle(precision):
    """BASELINE config[1] at full size: one training step of the CIFAR CNN at batch 128."""
    x, y = workloads.synthetic_batch("cnn", 128)
    wl = workloads.cifar_cnn(seed=0)
    model = orc.build_model("cnn", seed=0)
    loss_val, grads = wl.train_step(sp.Adam(), x, y)
    loss0, grads0 = model.train_step(orc.AdamState(), x, y)
    assert_close(loss_val.array, loss0, LOSS_RTOL[precision], "loss")
    for i, (p, p0) in enumerate(zip(wl.learnables, model.params)):
        assert_grad_close(grads[p], grads0[p0], precision, f"gradient {i}")
        assert_first_adam_step_close(p.array, p0.value, grads0[p0], WEIGHT_RTOL[precision],
                                     GRAD_RTOL[precision])


def test_vgg_small_one_step_vs_oracle(precision):
    """BASELINE config[3] family (SAME 3x3 convs, 64-256 channels) at a CPU-checkable size."""
    x, y = workloads.synthetic_batch("vgg", 4, image=16)
    widths = (16, 16, 32, 32, 64, 64)
    wl = workloads.vgg6(seed=0, image=16, widths=widths)
    model = orc.build_model("vgg", seed=0, image=16, widths=widths)
    loss_val, grads = wl.train_step(sp.Adam(), x, y)
    loss0, grads0 = model.train_step(orc.AdamState(), x, y)
    assert_close(loss_val.array, loss0, LOSS_RTOL[precision], "loss")
    for i, (p, p0) in enumerate(zip(wl.learnables, model.params)):
        assert_grad_close(grads[p], grads0[p0], precision, f"gradient {i}")


def test_pinned_host_batches_upload_without_a_host_pass():
    """float64 / float32 batches that already live in page-locked memory are DMA'd as they are
    and narrowed on the device (array.py from_host); values match the snapshot path."""
    import torch

    from smallpebble_b200.array import transfer_counters

    rng = np.random.default_rng(4)
    for dtype, tdtype in ((np.float64, torch.float64), (np.float32, torch.float32)):
        pinned = torch.empty((64, 32, 32, 3), dtype=tdtype).pin_memory().numpy()
        pinned[...] = rng.random(pinned.shape)
        before = transfer_counters()["h2d_bytes"]
        v = sp.Variable(pinned)
        assert v.array._host is pinned  # no staging copy was made
        got = np.asarray(sp.leaky_relu(v).array)
        assert transfer_counters()["h2d_bytes"] - before == pinned.nbytes
        assert np.array_equal(got, pinned.astype(np.float32))
        pageable = np.array(pinned)
        assert np.array_equal(np.asarray(sp.Variable(pageable).array), pinned.astype(np.float32))


def test_host_pipeline_narrows_float64_batches_exactly():
    """Page-locked float64 batches of >= 1 MB are converted to fp32 by the host thread pool while
    their copy is in flight (sp_h2d_f64_as_f32_begin / _finish): every value equals NumPy's
    round-to-nearest narrowing -- denormals, huge and tiny magnitudes, odd sizes, batches wrapped
    but never used (their staging slot is recycled), back-to-back uploads."""
    import ctypes

    import torch

    from smallpebble_b200 import _abi
    from smallpebble_b200 import array as A

    threads = ctypes.c_int(0)
    _abi.f.sp_host_pipe_threads(threads)
    assert threads.value >= 1
    rng = np.random.default_rng(8)
    sizes = [131072, 131072 + 5, 393216, 1_000_003, 65536 * 3 + 4095]
    batches = []
    for n in sizes:
        x = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
        x[...] = (rng.random(n) - 0.5) * 10.0 ** rng.integers(-44, 38, n)
        x[:7] = [0.0, -0.0, 1e-45, -3e-39, 3.4028235e38, 1.0 + 2.0 ** -24, 1.0 + 2.0 ** -24 + 2.0 ** -40]
        batches.append(x)
    with np.errstate(over="ignore", under="ignore"):
        want = [x.astype(np.float32) for x in batches]
    default = A.host_pipe  # (off: measured slower than the plain float64 copy on this pool's hosts)
    A.host_pipe = True
    try:
        before = A.transfer_counters()["h2d_bytes"]
        for rep in range(6):
            held = [sp.Variable(x) for x in batches]          # five uploads started, three slots
            unused = [sp.Variable(x) for x in batches[:2]]    # wrapped, never read
            for v, w in zip(held, want):
                v.array.ptr                                    # on the device now
                got = np.asarray(v.array)
                assert np.array_equal(got.view(np.uint32), w.view(np.uint32))
            del unused
        moved = A.transfer_counters()["h2d_bytes"] - before
        assert moved == 6 * sum(4 * n for n in sizes)          # fp32 on the wire, used arrays only
        A.host_pipe = False
        for x, w in zip(batches, want):                        # the device-narrowing path: same bits
            v = sp.Variable(x)
            v.array.ptr
            assert np.array_equal(np.asarray(v.array).view(np.uint32), w.view(np.uint32))
    finally:
        A.host_pipe = default


def test_device_array_free_list_never_aliases_live_buffers():
    """array.py recycles buffers by byte size; a buffer must not come back while a view or an
    in-place-rebound alias of it is alive, and it must come back once the last holder dies."""
    from smallpebble_b200.array import DeviceArray, trim_pool

    trim_pool()
    a = DeviceArray.full((1024,), 1.0)
    pa = a.ptr
    view = a.reshape(32, 32)          # shares a's buffer
    del a
    b = DeviceArray.full((1024,), 2.0)  # same size: must NOT reuse the viewed buffer
    assert b.ptr != pa
    assert np.all(np.asarray(view) == 1.0)
    del view                           # last holder: the buffer goes back to the free list
    c = DeviceArray.empty((1024,))
    assert c.ptr == pa
    # in-place arithmetic rebinding keeps exactly one owner
    x = sp.Variable(np.ones((64, 64))).array
    x.ptr
    y = x
    y += 1.0
    assert np.all(np.asarray(y) == 2.0)
    del b, c, x, y
    trim_pool()


def test_batch_yields_pinned_minibatches_with_reference_values():
    """sp.batch (sp.py:586-593): same RNG draws and values as the reference; on a GPU box the
    batch lives in page-locked memory so that Variable(xbatch) is a direct DMA."""
    from smallpebble_b200.array import _is_pinned

    rng = np.random.default_rng(3)
    X, y = rng.random((500, 32, 32, 3)), rng.integers(0, 10, 500)
    np.random.seed(7)   # what the reference's generator draws from the global RNG
    idx = np.random.randint(0, y.size, 128)
    idx2 = np.random.randint(0, y.size, 128)
    gen = sp.batch(X, y, 128, seed=7)
    xb, yb = next(gen)
    assert np.array_equal(xb, X[idx]) and np.array_equal(yb, y[idx])
    assert xb.dtype == np.float64 and _is_pinned(xb)
    xb2, _ = next(gen)
    assert np.array_equal(xb2, X[idx2]) and xb2 is not xb   # a fresh array per batch
    v = sp.Variable(xb)
    assert v.array._host is xb
    assert np.array_equal(np.asarray(v.array), X[idx].astype(np.float32))
    small = next(sp.batch(np.arange(20.0).reshape(10, 2), np.arange(10), 4, seed=1))[0]
    assert small.shape == (4, 2)   # below the pinning threshold: plain fancy indexing


def test_programmatic_dependent_launch_changes_no_bits():
    """Kernels are chained with programmatic dependent launch (griddepcontrol: the next kernel's
    launch and prologue overlap the current one, common.cuh).  Five Adam steps of the CIFAR CNN
    with it on and off (policy bit 9) must leave bit-identical weights and losses: the overlap
    never lets a kernel read or overwrite anything before its producer / last reader is done."""
    from smallpebble_b200 import core

    x, y = workloads.synthetic_batch("cnn", 32)

    def run(policy):
        core.debug_set_kernel_policy(policy)
        try:
            wl = workloads.cifar_cnn(seed=0)
            adam = sp.Adam()
            losses = []
            for _ in range(5):
                loss_val, _ = wl.train_step(adam, x, y)
                losses.append(np.asarray(loss_val.array))
            return losses, [np.asarray(p.array) for p in wl.learnables]
        finally:
            core.debug_set_kernel_policy(0)

    losses_on, weights_on = run(0)
    losses_off, weights_off = run(512)
    for a, b in zip(losses_on, losses_off):
        assert np.array_equal(a, b)
    for a, b in zip(weights_on, weights_off):
        assert np.array_equal(a, b)


def test_softmax_is_folded_into_cross_entropy_and_still_readable():
    """softmax(op_result) is a deferred array: cross_entropy's kernel fills it (one launch for
    probabilities + loss); reading it first launches the ordinary kernel.  Same bits both ways."""
    from smallpebble_b200 import _abi

    rng = np.random.default_rng(3)
    logits = rng.random((64, 10)) * 6 - 3
    labels = rng.integers(0, 10, 64)
    w = sp.Variable(np.eye(10))

    def run(read_first):
        z = sp.matmul(sp.Variable(logits), w)   # an op result feeds softmax
        p = sp.softmax(z)
        assert p.array.pending_tag is not None
        probs_early = np.asarray(p.array) if read_first else None
        before = _abi.counter.launches
        loss = sp.cross_entropy(p, labels)
        launched = _abi.counter.launches - before
        assert p.array.pending_tag is None
        g = sp.get_gradients(loss)
        return np.asarray(p.array), np.asarray(loss.array), np.asarray(g[z]), probs_early, launched

    p1, l1, g1, _, n1 = run(False)
    p2, l2, g2, early, _ = run(True)
    assert n1 == 1                       # probabilities and loss came out of one launch
    assert np.array_equal(p1, p2) and np.array_equal(p2, early)
    assert np.array_equal(l1, l2) and np.array_equal(g1, g2)
    ref = np.exp(logits - logits.max(1, keepdims=True))
    ref /= ref.sum(1, keepdims=True)
    assert_close(p1, ref, 1e-5, "probs")


@pytest.mark.parametrize("rows,k,n,bias_shape", [(128, 1152, 10, (10,)), (37, 300, 16, (1, 16)),
                                                  (256, 4608, 10, (10,))])
def test_classifier_bias_fusions_match_the_separate_ops(rows, k, n, bias_shape):
    """matmul(op_result, w) + bias for N <= 16 is ONE kernel (deferred matmul consumed by add),
    and the bias gradient comes out of the softmax-cross-entropy backward kernel as a
    by-product.  Values and the weight / input gradients are bit-identical to the separate
    launches; the bias gradient is the same sum in a different (fixed) order."""
    from smallpebble_b200 import _abi

    rng = np.random.default_rng(rows + k)
    x = rng.random((rows, k)) - 0.5
    w = (rng.random((k, n)) - 0.5) / np.sqrt(k)
    b = rng.random(bias_shape) - 0.5
    labels = rng.integers(0, n, rows)

    from smallpebble_b200 import core

    def run(fuse):
        core._fuse_bias = fuse  # only these two fusions: the loss kernels stay the same
        core._fuse_head = False  # (the whole-head kernel has its own test below)
        try:
            xv, wv, bv = sp.Variable(x), sp.Variable(w), sp.Variable(b)
            pre = sp.mul(xv, sp.Variable(np.full(x.shape, 1.25)))  # an op result feeds matmul
            before = _abi.counter.launches
            z = sp.matmul(pre, wv)
            logits = sp.add(z, bv)
            fwd_launches = _abi.counter.launches - before
            assert (z.array.pending_tag is not None) == fuse  # never launched when fused
            loss = sp.cross_entropy(sp.softmax(logits), labels)
            g = sp.get_gradients(loss)
            out = [np.asarray(logits.array), np.asarray(loss.array), np.asarray(g[wv]),
                   np.asarray(g[xv]), np.asarray(g[bv]), np.asarray(z.array)]
            return out, fwd_launches
        finally:
            core._fuse_bias = True
            core._fuse_head = True

    fused, n_fused = run(True)
    plain, n_plain = run(False)
    assert n_fused == 1 and n_plain == 2
    for i in (0, 1, 2, 3, 5):
        assert np.array_equal(fused[i], plain[i]), i
    assert fused[4].shape == bias_shape
    assert_close(fused[4], plain[4], 1e-6, "bias gradient")
    ref = (x * 1.25) @ w
    assert_close(fused[5], ref, 2e-5 if n < 16 else 2e-3, "deferred matmul value, read after the fact")


@pytest.mark.parametrize("precision", ["tf32", "tf32x1", "fp32"])
@pytest.mark.parametrize("rows,k,n", [(128, 1152, 10), (200, 100, 10), (37, 300, 16), (256, 16384, 10),
                                      (64, 512, 3), (1500, 64, 12)])
def test_classifier_head_in_one_kernel_is_bit_identical(rows, k, n, precision):
    """cross_entropy(softmax(matmul(h, w) + b), y) for N <= 16 is ONE kernel
    (sp_gemm_bias_softmax_xent: the last block of the skinny GEMM runs the softmax, the loss
    sum, p - onehot and its column sums).  Logits, probabilities, loss and every gradient equal
    the separate launches bit for bit -- repeatedly (the arrival counter resets itself)."""
    from smallpebble_b200 import _abi, core

    rng = np.random.default_rng(rows + k + n)
    x = rng.random((rows, k)) - 0.5
    w = (rng.random((k, n)) - 0.5) * 4 / np.sqrt(k)
    b = rng.random((n,)) - 0.5
    labels = rng.integers(0, n, rows)
    sp.set_precision(precision)

    def run(fuse):
        core._fuse_head = fuse
        try:
            xv, wv, bv = sp.Variable(x), sp.Variable(w), sp.Variable(b)
            pre = sp.mul(xv, sp.Variable(np.full(x.shape, 1.25)))  # an op result feeds matmul
            before = _abi.counter.launches
            logits = sp.add(sp.matmul(pre, wv), bv)
            probs = sp.softmax(logits)
            loss = sp.cross_entropy(probs, labels)
            fwd_launches = _abi.counter.launches - before
            g = sp.get_gradients(loss)
            out = [np.asarray(logits.array), np.asarray(probs.array), np.asarray(loss.array),
                   np.asarray(g[wv]), np.asarray(g[xv]), np.asarray(g[bv])]
            return out, fwd_launches
        finally:
            core._fuse_head = True

    try:
        fused, n_fused = run(True)
        again, _ = run(True)
        plain, n_plain = run(False)
    finally:
        sp.set_precision("tf32")
    assert n_plain == 2
    assert n_fused == 1
    for a, c, d in zip(fused, again, plain):
        assert np.array_equal(a, d) and np.array_equal(c, d)
    z = (x * 1.25) @ w + b
    e = np.exp(z - z.max(1, keepdims=True))
    ref = -np.log(e[np.arange(rows), labels] / e.sum(1)).sum()
    assert_close(fused[2], ref, 2e-5 if precision != "tf32x1" else 2e-3, "loss")


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
@pytest.mark.parametrize("rows,k,n", [(200, 784, 100), (200, 100, 100), (37, 300, 48), (128, 2048, 64),
                                      (5, 64, 17), (64, 70, 130)])
def test_hidden_linear_layer_in_one_kernel_is_bit_identical(rows, k, n, precision):
    """leaky_relu(matmul(x, w) + b) of a small layer (the README MLP's hidden layers) is ONE
    launch (sp_gemm_bias_leaky: K split over a thread-block cluster, partial tiles added through
    distributed shared memory in rank order, bias and activation in the epilogue).  Values and
    gradients equal the separate launches bit for bit, and the float64 product within fp32."""
    from smallpebble_b200 import _abi, core

    rng = np.random.default_rng(rows + k + n)
    x = rng.random((rows, k)) - 0.5
    w = (rng.random((k, n)) - 0.5) * 4 / np.sqrt(k)
    b = rng.random((n,)) - 0.5
    sp.set_precision(precision)

    def run(fuse):
        core._fuse_linear = fuse
        try:
            xv, wv, bv = sp.Variable(x), sp.Variable(w), sp.Variable(b)
            before = _abi.counter.launches
            z = sp.add(sp.matmul(xv, wv), bv)
            h = sp.leaky_relu(z, 0.05)
            launches = _abi.counter.launches - before
            gy = np.random.default_rng(9).random(h.shape) - 0.5
            g = sp.get_gradients(sp.mul(h, sp.Variable(gy)))
            return [np.asarray(h.array), np.asarray(g[wv]), np.asarray(g[xv]), np.asarray(g[bv]),
                    np.asarray(z.array)], launches
        finally:
            core._fuse_linear = True

    try:
        fused, n_fused = run(True)
        plain, n_plain = run(False)
    finally:
        sp.set_precision("tf32")
    if precision == "tf32":  # (compensated mode: small forward products run exactly on CUDA cores)
        assert n_fused == 1 and n_plain >= 3
    for a, c in zip(fused, plain):
        assert np.array_equal(a, c)
    z = x @ w + b
    assert_close(fused[0], np.where(z > 0, z, 0.05 * z), 2e-5, "layer output")
    assert_close(fused[4], z, 2e-5, "pre-activation, launched on demand")


@pytest.mark.parametrize("policy", [0, 32 + 256, 32 + 128], ids=["policy", "tma2", "halo"])
@pytest.mark.parametrize("alpha", [0.02, 0.0])
def test_conv_leaky_epilogue_fusion_is_bit_identical(policy, alpha):
    """leaky_relu(conv2d(...)) feeding another conv runs as ONE kernel (activation in the conv
    epilogue: the 2-SM kernels, the first-layer tensor-core kernel, or conv + in-place leaky for
    the others) and the pre-activation tensor is never stored; leaky's gradient then takes the
    sign from the activation's output.  Everything must equal the separate launches bit for
    bit; alpha = 0 (sign no longer preserved) must not take the fused path."""
    from smallpebble_b200 import core

    rng = np.random.default_rng(11)
    x = rng.random((4, 12, 12, 3)) - 0.5
    w1 = (rng.random((3, 3, 3, 32)) - 0.5) * 0.4
    w2 = (rng.random((3, 3, 32, 64)) - 0.5) * 0.1
    w3 = (rng.random((3, 3, 64, 64)) - 0.5) * 0.1
    seed = rng.random((4, 6, 6, 64)) - 0.5

    def run(fuse):
        old = sp.get_precision()
        sp.set_precision("tf32")
        core._fuse_conv_leaky = fuse
        core.debug_set_kernel_policy(policy)
        try:
            xv, v1, v2, v3 = (sp.Variable(a) for a in (x, w1, w2, w3))
            c1 = sp.conv2d(xv, v1, "SAME")
            a1 = sp.leaky_relu(c1, alpha)               # -> conv: epilogue fusion
            c2 = sp.conv2d(a1, v2, "SAME")
            a2 = sp.leaky_relu(c2, alpha)               # -> pool: pool fusion, conv stored
            p2 = sp.maxpool2d(a2, 2, 2, strides=[2, 2])
            c3 = sp.conv2d(p2, v3, "SAME")
            a3 = sp.leaky_relu(c3, alpha)               # -> mul: plain materialisation
            out = sp.mul(a3, sp.Variable(seed))
            g = sp.get_gradients(out)
            fused_c1 = c1.array.pending_tag is not None  # never launched on its own
            res = [np.asarray(t) for t in (out.array, a1.array, a2.array, a3.array, g[xv], g[v1],
                                           g[v2], g[v3], g[c1], g[a1], c1.array)]
            return res, fused_c1
        finally:
            core.debug_set_kernel_policy(0)
            core._fuse_conv_leaky = True
            sp.set_precision(old)

    fused, was_fused = run(True)
    plain, was_plain = run(False)
    assert was_fused == (alpha > 0) and not was_plain
    for i, (a, b) in enumerate(zip(fused, plain)):
        assert np.array_equal(a, b), i


def test_adam_follows_a_rebound_parameter_array():
    """Adam caches the parameter / moment pointer tables between steps; rebinding
    `variable.array` (what the reference's own optimiser does every step, sp.py:583, and what a
    user does to load weights) must invalidate that cache: a run that swaps every parameter for
    a copy of itself half-way must end with exactly the same weights as one that does not."""
    x, y = workloads.synthetic_batch("mlp", 64)

    def run(rebind):
        wl = workloads.mnist_mlp(seed=0)
        adam = sp.Adam()
        for step in range(4):
            if rebind and step == 2:
                for p in wl.learnables:
                    p.array = np.asarray(p.array).copy()
            wl.train_step(adam, x, y)
        return [np.asarray(p.array) for p in wl.learnables], [adam.t[p] for p in wl.learnables]

    w_plain, t_plain = run(False)
    w_rebound, t_rebound = run(True)
    assert t_plain == t_rebound == [4] * len(t_plain)
    for a, b in zip(w_plain, w_rebound):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("optimiser", ["adam", "sgd"])
def test_leaf_gradient_read_after_the_update_sees_the_old_weights(optimiser):
    """`get_gradients` defers the gradient of non-learnable leaves (the image batch) until it
    is read.  The reference evaluates every gradient inside get_gradients (sp.py:77-87) and
    Adam rebinds `variable.array` (sp.py:583): reading `g[x]` AFTER the optimiser step must give
    the bits it would have given before the step."""
    x, y = workloads.synthetic_batch("cnn", 8)

    def run(read_before):
        wl = workloads.cifar_cnn(seed=0)
        xv = sp.Variable(x)
        wl.x_in.assign_value(xv)
        wl.y_true.assign_value(y)
        loss = wl.loss.run()
        g = sp.get_gradients(loss)
        before = np.asarray(g[xv]).copy() if read_before else None
        if optimiser == "adam":
            sp.Adam().training_step(wl.learnables, g)
        else:
            sp.sgd_step(wl.learnables, g, 0.1)
        after = np.asarray(g[xv]).copy()
        return before, after

    before, after = run(True)
    _, late = run(False)
    assert np.abs(before).max() > 0
    assert np.array_equal(before, after)
    assert np.array_equal(before, late)


@pytest.mark.parametrize("precision", ["tf32", "tf32x1"])
@pytest.mark.parametrize("batch,image,widths,pads", [
    (32, 32, (32, 128, 128), ("SAME", "VALID", "VALID")),    # README CNN: split-K cluster epilogue
    (140, 32, (32, 64, 32), ("SAME", "VALID", "VALID")),     # too few k stages to split: direct epilogue
    (96, 30, (32, 64, 64), ("SAME", "VALID", "SAME")),       # 13x13 -> 7x7 pool: clipped windows
])
def test_conv_data_gradient_scattered_through_the_pool_is_bit_identical(batch, image, widths, pads,
                                                                        precision):
    """conv -> leaky_relu -> maxpool2d(2, 2, stride 2) -> conv: the second conv's data gradient
    is scattered through the pool's argmax and multiplied by leaky_relu's slope in the epilogue
    of the conv kernel (sp_conv2d_dgrad_unpool_leaky) -- no pooled gradient tensor, no pool
    backward launch -- and every gradient equals the separate launches bit for bit."""
    from smallpebble_b200 import _abi, core

    rng = np.random.default_rng(batch + image)
    x = rng.random((batch, image, image, 3)) - 0.5
    ws, cin = [], 3
    for wd in widths:
        ws.append((rng.random((3, 3, cin, wd)) - 0.5) * 2.0 / np.sqrt(9 * cin))
        cin = wd
    sp.set_precision(precision)
    calls = []
    orig = _abi.f.sp_conv2d_dgrad_unpool_leaky

    def spy(*args):
        calls.append(1)
        return orig(*args)

    def run():
        xv = sp.Variable(x)
        wv = [sp.learnable(sp.Variable(w)) for w in ws]
        h = xv
        acts = []
        for i, (w, pad) in enumerate(zip(wv, pads)):
            h = sp.conv2d(h, w, pad)
            acts.append(h)
            if i < len(wv) - 1:
                h = sp.maxpool2d(sp.leaky_relu(h, 0.1), 2, 2, strides=[2, 2])
        gy = np.random.default_rng(5).random(h.shape) - 0.5
        g = sp.get_gradients(sp.mul(h, sp.Variable(gy)))
        return [np.asarray(g[w]) for w in wv] + [np.asarray(g[acts[1]]), np.asarray(g[xv])]

    _abi.f.sp_conv2d_dgrad_unpool_leaky = spy
    default = core._fuse_dgrad_unpool  # (off: measured slower inside the backward graph)
    try:
        core._fuse_dgrad_unpool = True
        fused = run()
        n_fused = len(calls)
        core._fuse_dgrad_unpool = False
        plain = run()
    finally:
        _abi.f.sp_conv2d_dgrad_unpool_leaky = orig
        core._fuse_dgrad_unpool = default
        sp.set_precision("tf32")
    assert n_fused == 1 and len(calls) == 1  # the third conv's data gradient; never when off
    for a, b in zip(fused, plain):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("precision", ["tf32", "tf32x1"])
@pytest.mark.parametrize("batch,image,widths,pads", [
    (128, 32, (32, 128, 128), ("VALID", "VALID", "VALID")),  # the README CNN: 13x13 -> 7x7, 5x5 -> 3x3
    (16, 32, (32, 128, 128), ("VALID", "VALID", "VALID")),   # few tiles: every layer on split-K clusters
    (40, 32, (32, 64, 256), ("SAME", "SAME", "SAME")),        # 16x16 -> 8x8 (two rows per tile), 256 filters
    (24, 48, (32, 64, 64), ("SAME", "VALID", "SAME")),        # 22x22 -> 11x11, 11x11 -> 6x6
])
def test_tcgen05_conv_leaky_pool_in_one_kernel_is_bit_identical(batch, image, widths, pads, precision):
    """conv2d -> leaky_relu -> maxpool2d(2, 2, stride 2) of a tensor-core layer as ONE kernel: the
    2-SM tcgen05 kernel's CTA tiles are laid over whole pooling windows and its epilogue pools
    (sp_conv2d_leaky_pool2_fwd; conv_tma2.cu EPI = 3).  Pooled values, argmax bytes, the (r, l)
    split the next compensated conv consumes, every gradient -- including the input gradient,
    whose pool backward takes leaky's slope from the pooled output -- and the full-resolution
    tensor launched on demand equal the separate launches bit for bit."""
    from smallpebble_b200 import _abi, core

    rng = np.random.default_rng(batch + image)
    x = rng.random((batch, image, image, 3)) - 0.5
    ws, cin = [], 3
    for wd in widths:
        ws.append((rng.random((3, 3, cin, wd)) - 0.5) * 2.0 / np.sqrt(9 * cin))
        cin = wd
    sp.set_precision(precision)
    calls = []
    orig = _abi.f.sp_conv2d_leaky_pool2_fwd

    def spy(*args):
        calls.append(args[5].C if hasattr(args[5], "C") else 0)
        return orig(*args)

    def run(read_full):
        xv = sp.Variable(x)
        wv = [sp.learnable(sp.Variable(w)) for w in ws]
        h, convs, pools = xv, [], []
        for w, pad in zip(wv, pads):
            h = sp.conv2d(h, w, pad)
            convs.append(h)
            h = sp.maxpool2d(sp.leaky_relu(h, 0.1), 2, 2, strides=[2, 2])
            pools.append(h)
        gy = np.random.default_rng(5).random(h.shape) - 0.5
        g = sp.get_gradients(sp.mul(h, sp.Variable(gy)))
        res = [np.asarray(p.array) for p in pools] + [np.asarray(p._argmax) for p in pools]
        res += [np.asarray(g[w]) for w in wv] + [np.asarray(g[xv])]
        if read_full:
            res += [np.asarray(c.array) for c in convs]  # launched on demand
        return res

    _abi.f.sp_conv2d_leaky_pool2_fwd = spy
    try:
        core._fuse_conv_pool = True
        run(False)                       # (first pass: the consumers' split hints get known)
        del calls[:]
        fused = run(True)
        n_fused = len(calls)
        core._fuse_conv_pool = False
        plain = run(True)
    finally:
        _abi.f.sp_conv2d_leaky_pool2_fwd = orig
        core._fuse_conv_pool = True
        sp.set_precision("tf32")
    assert len(calls) == n_fused and n_fused >= 1, calls
    if batch >= 24:  # (below that the layers are too small for the 2-SM kernel's policy)
        assert n_fused >= 2 and any(c >= 32 for c in calls), calls  # tensor-core layers too
    for i, (a, b) in enumerate(zip(fused, plain)):
        assert np.array_equal(a, b), i


@pytest.mark.parametrize("precision", ["tf32", "tf32x1"])
@pytest.mark.parametrize("shape,k,padding", [((16, 32, 32, 3), 32, "SAME"), ((5, 14, 18, 3), 16, "VALID"),
                                             ((3, 40, 36, 3), 32, "SAME"), ((2, 10, 10, 3), 64, "VALID")])
def test_first_layer_conv_leaky_pool_in_one_kernel_is_bit_identical(shape, k, padding, precision):
    """conv2d (3x3, C = 3) -> leaky_relu -> maxpool2d(2, 2, stride 2) as ONE forward kernel
    (sp_conv2d_leaky_pool2_fwd): pooled values, argmax bytes, the (r, l) split the next
    compensated conv2d consumes and every gradient equal the separate launches bit for bit; the
    64-filter case has no such kernel and must take the old path."""
    from smallpebble_b200 import _abi, core

    rng = np.random.default_rng(21)
    x = rng.random(shape) - 0.5
    w = (rng.random((3, 3, 3, k)) - 0.5) * 0.5
    w2 = (rng.random((3, 3, k, 128)) - 0.5) * 0.2  # (shape0: the README CNN's second conv)
    sp.set_precision(precision)
    calls = []
    orig = _abi.f.sp_conv2d_leaky_pool2_fwd

    def spy(*args):
        calls.append(args[4])  # the split pointer (None the first time: no hint yet)
        return orig(*args)

    def run(read_dx):
        xv, wv, w2v = sp.Variable(x), sp.Variable(w), sp.Variable(w2)
        sp.learnable(wv), sp.learnable(w2v)
        h = sp.conv2d(xv, wv, padding)
        p = sp.maxpool2d(sp.leaky_relu(h, 0.1), 2, 2, strides=[2, 2])
        out = sp.conv2d(p, w2v, "VALID")
        gy = np.random.default_rng(22).random(out.shape) - 0.5
        g = sp.get_gradients(sp.mul(out, sp.Variable(gy)))
        res = [np.asarray(p.array), np.asarray(p._argmax), np.asarray(out.array),
               np.asarray(g[wv]), np.asarray(g[w2v])]
        if read_dx:
            res.append(np.asarray(g[xv]))
            res.append(np.asarray(h.array))  # the full-resolution tensor, launched on demand
        return res

    _abi.f.sp_conv2d_leaky_pool2_fwd = spy
    try:
        core._fuse_conv_pool = True
        fused_a = run(False)
        fused_b = run(True)   # second run: the split hint of the consumer conv is known
        n_fused = len(calls)
        core._fuse_conv_pool = False
        plain = run(True)
    finally:
        _abi.f.sp_conv2d_leaky_pool2_fwd = orig
        core._fuse_conv_pool = True
        sp.set_precision("tf32")
    assert len(calls) == n_fused
    if k <= 32:
        # (a pooled tensor whose consumer wants another split layout is pooled the old way)
        assert n_fused >= 1
        if precision == "tf32" and shape[0] == 16:
            assert n_fused == 2 and calls[1] is not None  # the split comes out of the same kernel
    else:
        assert n_fused == 0
    for a, b in zip(fused_a, plain):
        assert np.array_equal(a, b)
    for a, b in zip(fused_b, plain):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("shape,k", [((16, 32, 32, 3), 32), ((5, 14, 18, 3), 16), ((3, 10, 10, 3), 64)])
def test_first_layer_filter_gradient_from_the_pooled_gradient_is_bit_identical(shape, k):
    """conv2d (3x3, C = 3) -> leaky_relu -> maxpool2d(2, 2, stride 2): the filter gradient expands
    the POOLED gradient inside its own kernel (sp_conv2d_wgrad_pooled) instead of reading the
    full-resolution gradient sp_maxpool2d_leaky_bwd would write -- the same products (summed
    over differently shaped tiles), and the image gradient, which needs the full-resolution
    tensor, still comes out bit-identical."""
    from smallpebble_b200 import _abi

    rng = np.random.default_rng(11)
    x = rng.random(shape) - 0.5
    w = (rng.random((3, 3, 3, k)) - 0.5) * 0.5
    sp.set_precision("tf32")
    calls = []
    orig = _abi.f.sp_conv2d_wgrad_pooled

    def spy(*args):
        calls.append(1)
        return orig(*args)

    def run(fuse, read_dx):
        sp.set_fusion(fuse)
        xv, wv = sp.Variable(x), sp.Variable(w)
        sp.learnable(wv)
        h = sp.conv2d(xv, wv, "VALID")
        h = sp.maxpool2d(sp.leaky_relu(h, 0.1), 2, 2, strides=[2, 2])
        gy = np.random.default_rng(12).random(h.shape) - 0.5
        g = sp.get_gradients(sp.mul(h, sp.Variable(gy)))
        dw = np.asarray(g[wv])
        dx = np.asarray(g[xv]) if read_dx else None
        return np.asarray(h.array), dw, dx

    _abi.f.sp_conv2d_wgrad_pooled = spy
    try:
        y1, dw1, dx1 = run(True, True)
        n_fused = len(calls)
        y0, dw0, dx0 = run(False, True)
    finally:
        _abi.f.sp_conv2d_wgrad_pooled = orig
        sp.set_fusion(True)
    assert n_fused == 1 and len(calls) == 1  # taken when fusing, never otherwise
    assert np.array_equal(y1, y0) and np.array_equal(dx1, dx0)
    # same products, but tiles of an even number of rows: the fp32 summation order differs
    assert_close(dw1, dw0, 2e-6, "fused vs separate launches")
    # and against the float64 oracle
    xo, wo = orc.Node(x), orc.Node(w)
    ho = orc.op_maxpool2d(orc.op_leaky_relu(orc.op_conv2d(xo, wo, "VALID"), 0.1), 2, 2, "SAME", (2, 2))
    gy = np.random.default_rng(12).random(ho.value.shape) - 0.5
    go = orc.backprop(orc.op_mul(ho, orc.Node(gy)))
    assert_close(dw1, go[wo], 2e-3, "dW vs oracle")
    assert_close(dx1, go[xo], 2e-3, "dX vs oracle")
