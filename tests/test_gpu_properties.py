"""GPU: size-independent properties at BASELINE.json's full sizes (where the float64 oracle
would take minutes): adjointness of the three conv kernels, linearity, pool round-trips,
determinism, agreement of the two kernel families."""

import numpy as np
import pytest

import smallpebble_b200 as sp
from conftest import RTOL, assert_close, scaled_err
from smallpebble_b200 import _abi, core

pytestmark = pytest.mark.gpu


def dot(a, b):
    return float(np.sum(np.asarray(a, np.float64) * np.asarray(b, np.float64)))


CONFIGS = [
    # name, x shape, w shape, strides, padding
    ("cnn_conv1_b128", (128, 32, 32, 3), (3, 3, 3, 32), (1, 1), "VALID"),
    ("cnn_conv2_b128", (128, 15, 15, 32), (3, 3, 32, 128), (1, 1), "VALID"),
    ("cnn_conv3_b128", (128, 7, 7, 128), (3, 3, 128, 128), (1, 1), "VALID"),
    ("vgg_conv_64_64", (32, 64, 64, 64), (3, 3, 64, 64), (1, 1), "SAME"),
    ("sweep_k5_s2", (64, 32, 32, 64), (5, 5, 64, 64), (2, 2), "SAME"),
    ("sweep_k1", (64, 32, 32, 128), (1, 1, 128, 256), (1, 1), "SAME"),
    # BASELINE configs[3] (VGG-style, batch 256): large enough for the 2-SM TMA kernels
    ("vgg_conv_128_256_b256", (256, 16, 16, 128), (3, 3, 128, 256), (1, 1), "SAME"),
    ("vgg_conv_256_256_b256", (256, 16, 16, 256), (3, 3, 256, 256), (1, 1), "SAME"),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_conv_adjoint_identities(cfg, precision):
    """<conv(x,w), gy> == <x, dgrad(gy)> == <w, wgrad(gy)>: the three kernels are consistent
    linear maps of each other (exact in real arithmetic because conv is bilinear)."""
    name, xs, ws, strides, pad = cfg
    rng = np.random.default_rng(7)
    x, w = rng.random(xs) - 0.5, rng.random(ws) - 0.5
    xv, wv = sp.Variable(x), sp.Variable(w)
    y = sp.conv2d(xv, wv, pad, strides)
    gy = rng.random(y.shape) - 0.5
    g = sp.get_gradients(sp.mul(y, sp.Variable(gy)))
    lhs = dot(y.array, gy)
    # compare on the scale of the accumulated magnitudes, not of the (cancelling) sum
    scale = float(np.sum(np.abs(np.asarray(y.array, np.float64) * gy)))
    tol = RTOL[precision]
    assert abs(lhs - dot(x, g[xv])) <= tol * scale
    assert abs(lhs - dot(w, g[wv])) <= tol * scale


@pytest.mark.parametrize("cfg", CONFIGS[:4], ids=[c[0] for c in CONFIGS[:4]])
def test_tensor_core_kernel_agrees_with_fp32_kernel(cfg):
    name, xs, ws, strides, pad = cfg
    rng = np.random.default_rng(11)
    x, w = rng.random(xs) - 0.5, rng.random(ws) - 0.5

    def run(prec):
        sp.set_precision(prec)
        xv, wv = sp.Variable(x), sp.Variable(w)
        y = sp.conv2d(xv, wv, pad, strides)
        g = sp.get_gradients(y)
        return np.asarray(y.array), np.asarray(g[xv]), np.asarray(g[wv])

    old = sp.get_precision()
    try:
        core.debug_set_kernel_policy(1)
        tf = run("tf32")
        core.debug_set_kernel_policy(0)
        fp = run("fp32")
    finally:
        core.debug_set_kernel_policy(0)
        sp.set_precision(old)
    for a, b, what in zip(tf, fp, ("y", "dX", "dW")):
        assert_close(a, b, RTOL["tf32"], f"{name} {what}: tcgen05 vs CUDA-core")


TMA2_CONFIGS = [CONFIGS[1], CONFIGS[2], CONFIGS[3], CONFIGS[4], CONFIGS[6]]


@pytest.mark.parametrize("policy", [32, 32 + 128], ids=["im2col", "halo"])
@pytest.mark.parametrize("cfg", TMA2_CONFIGS, ids=[c[0] for c in TMA2_CONFIGS])
def test_tma2_kernel_agrees_with_fp32_kernel(cfg, policy):
    """The 2-SM TMA kernels (forced: policy bit 5, halo-reuse variant: + bit 7) against the
    CUDA-core fp32 kernels at full BASELINE sizes -- including the strided k5 case whose dgrad
    stays on the 1-SM kernel."""
    name, xs, ws, strides, pad = cfg
    rng = np.random.default_rng(13)
    x, w = rng.random(xs) - 0.5, rng.random(ws) - 0.5

    def run(prec, gy=None):
        sp.set_precision(prec)
        xv, wv = sp.Variable(x), sp.Variable(w)
        y = sp.conv2d(xv, wv, pad, strides)
        if gy is None:
            gy = rng.random(y.shape) - 0.5
        g = sp.get_gradients(sp.mul(y, sp.Variable(gy)))
        return (np.asarray(y.array), np.asarray(g[xv]), np.asarray(g[wv])), gy

    old = sp.get_precision()
    try:
        core.debug_set_kernel_policy(policy)
        tf, gy = run("tf32")
        tf_again, _ = run("tf32", gy)
        core.debug_set_kernel_policy(0)
        fp, _ = run("fp32", gy)
    finally:
        core.debug_set_kernel_policy(0)
        sp.set_precision(old)
    for a, b, c, what in zip(tf, fp, tf_again, ("y", "dX", "dW")):
        assert_close(a, b, RTOL["tf32"], f"{name} {what}: 2-SM TMA vs CUDA-core")
        assert np.array_equal(a, c), f"{name} {what}: not bitwise repeatable"


def test_conv_linearity_and_determinism(precision):
    rng = np.random.default_rng(3)
    x1, x2 = rng.random((64, 30, 30, 32)) - 0.5, rng.random((64, 30, 30, 32)) - 0.5
    w = rng.random((3, 3, 32, 64)) - 0.5
    wv = sp.Variable(w)
    f = lambda x: np.asarray(sp.conv2d(sp.Variable(x), wv).array, np.float64)  # noqa: E731
    y1, y2, y12 = f(x1), f(x2), f(x1 + 2 * x2)
    assert scaled_err(y12, y1 + 2 * y2) <= 2 * RTOL[precision]
    assert np.array_equal(f(x1), y1.astype(np.float64))  # bitwise repeatable
    xv = sp.Variable(x1)
    gw = [np.asarray(sp.get_gradients(sp.conv2d(xv, wv))[wv]) for _ in range(2)]
    assert np.array_equal(gw[0], gw[1])  # split-K wgrad has a fixed summation order


def test_maxpool_properties_full_size():
    rng = np.random.default_rng(5)
    x = (rng.random((128, 30, 30, 32)) - 0.5).astype(np.float32)
    xv = sp.Variable(x)
    y = sp.maxpool2d(xv, 2, 2, strides=[2, 2])
    yy, idx = np.asarray(y.array), np.asarray(y._argmax)
    # value == the window element the index points at; index in range; max >= every element
    win = x.reshape(128, 15, 2, 15, 2, 32).transpose(0, 1, 3, 5, 2, 4).reshape(128, 15, 15, 32, 4)
    assert idx.max() <= 3
    assert np.array_equal(np.take_along_axis(win, idx[..., None].astype(np.int64), -1)[..., 0], yy)
    assert np.array_equal(yy, win.max(-1)) and np.array_equal(idx, win.argmax(-1))
    # backward routes each upstream value to exactly one input: sums are preserved
    gy = rng.random(yy.shape).astype(np.float32)
    dx = np.asarray(sp.get_gradients(sp.mul(y, sp.Variable(gy)))[xv])
    assert abs(dx.sum(dtype=np.float64) - gy.sum(dtype=np.float64)) < 1e-3
    assert np.count_nonzero(dx) <= gy.size
    # idempotence: pooling an already pooled (constant-window) image returns it
    up = np.repeat(np.repeat(yy, 2, 1), 2, 2)
    again = np.asarray(sp.maxpool2d(sp.Variable(up), 2, 2, strides=[2, 2]).array)
    assert np.array_equal(again, yy)


def test_matmul_full_size_vs_float64(precision):
    rng = np.random.default_rng(9)
    a, b = rng.random((1024, 1152)) - 0.5, rng.random((1152, 256)) - 0.5
    av, bv = sp.Variable(a), sp.Variable(b)
    y = sp.matmul(av, bv)
    assert_close(y.array, a @ b, RTOL[precision], "matmul 1024x1152x256")
    g = sp.get_gradients(y)
    ones = np.ones((1024, 256))
    assert_close(g[av], ones @ b.T, RTOL[precision], "dA")
    assert_close(g[bv], a.T @ ones, RTOL[precision], "dB")


@pytest.mark.parametrize("m,k,n", [(256, 16384, 10), (67, 4100, 13), (700, 5000, 3), (8, 9000, 10)])
def test_skinny_matmul_long_k(m, k, n):
    """N <= 16 classifier GEMMs with a long k axis: cluster split-K rows kernel (forward, dA
    goes to the tiled kernels) and the cols kernel (dB = A^T g), against float64."""
    rng = np.random.default_rng(m + k + n)
    a, b = rng.random((m, k)) - 0.5, rng.random((k, n)) - 0.5
    av, bv = sp.Variable(a), sp.Variable(b)
    y = sp.matmul(av, bv)
    assert_close(y.array, a @ b, 2e-5, f"skinny {m}x{k}x{n}")
    gy = rng.random((m, n)) - 0.5
    g = sp.get_gradients(sp.mul(y, sp.Variable(gy)))
    assert_close(g[bv], a.T @ gy, 2e-5, "skinny dB")
    assert_close(g[av], gy @ b.T, 2e-3, "skinny dA")
    # deterministic: a second evaluation is bit-identical
    y2 = sp.matmul(av, bv)
    assert np.array_equal(np.asarray(y.array), np.asarray(y2.array))


@pytest.mark.parametrize("m,k,n", [(2048, 512, 256), (1000, 1024, 192), (4096, 256, 512)])
def test_matmul_on_the_two_sm_tma_kernels(m, k, n):
    """Dense GEMMs with whole 32 / 64-column blocks and enough work run as 1 x 1 convolutions on
    the 2-SM TMA-fed tcgen05 kernels (NN = fprop, NT = dgrad, TN = wgrad; contract_api.cu):
    forward and both gradients against float64, in the single-pass TF32 mode that takes them."""
    sp.set_precision("tf32x1")
    try:
        rng = np.random.default_rng(m + k + n)
        a, b = rng.random((m, k)) - 0.5, rng.random((k, n)) - 0.5
        av, bv = sp.Variable(a), sp.Variable(b)
        y = sp.matmul(av, bv)
        assert_close(y.array, a @ b, 2e-3, "y")
        gy = rng.random((m, n)) - 0.5
        g = sp.get_gradients(sp.mul(y, sp.Variable(gy)))
        assert_close(g[av], gy @ b.T, 2e-3, "dA")
        assert_close(g[bv], a.T @ gy, 2e-3, "dB")
    finally:
        sp.set_precision("tf32")


def test_elementwise_checksums_full_size():
    rng = np.random.default_rng(2)
    x = rng.random((128, 30, 30, 32)) - 0.5
    xv = sp.Variable(x)
    y = sp.leaky_relu(xv)
    x32 = x.astype(np.float32)
    assert np.array_equal(np.asarray(y.array), np.where(x32 > 0, x32, x32 * np.float32(0.02)))
    total = sp.sum(y)
    assert abs(float(total.array) - float(np.sum(np.asarray(y.array), dtype=np.float64))) < 0.5
    assert_close(sp.sum(y, axis=(0, 1, 2)).array,
                 np.asarray(y.array, np.float64).sum((0, 1, 2)), 1e-5, "axis sum")


def test_empty_and_ragged_inputs():
    """Edge cases: empty batch, batch of one, window == image, stride > kernel."""
    w = sp.Variable(np.random.random((3, 3, 4, 8)))
    y = sp.conv2d(sp.Variable(np.zeros((0, 8, 8, 4))), w)
    assert y.shape == (0, 8, 8, 8)
    y = sp.conv2d(sp.Variable(np.random.random((1, 3, 3, 4))), w, "VALID")
    assert y.shape == (1, 1, 1, 8)
    p = sp.maxpool2d(sp.Variable(np.random.random((2, 7, 7, 3))), 2, 2, "VALID", [5, 5])
    assert p.shape == (2, 2, 2, 3)
    s = sp.sum(sp.Variable(np.zeros((0, 5))), axis=0)
    assert np.array_equal(np.asarray(s.array), np.zeros(5))
    m = sp.matmul(sp.Variable(np.zeros((4, 0))), sp.Variable(np.zeros((0, 3))))
    assert np.array_equal(np.asarray(m.array), np.zeros((4, 3)))


def _random_geometries(seed, count):
    """Random conv geometries whose channel counts make them eligible for the 2-SM kernels:
    odd image sizes, windows up to 5x4, both paddings, strides 1-3, tiny and ragged batches."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        n = int(rng.choice([1, 2, 3, 5]))
        h, w = int(rng.integers(1, 20)), int(rng.integers(1, 20))
        kh, kw = int(rng.integers(1, 6)), int(rng.integers(1, 5))
        c = int(rng.choice([32, 64, 96]))
        k = int(rng.choice([64, 128, 192]))
        pad = str(rng.choice(["SAME", "VALID"]))
        s = (int(rng.integers(1, 4)), int(rng.integers(1, 4))) if rng.random() < 0.3 else (1, 1)
        if pad == "VALID" and (kh > h or kw > w):
            continue
        out.append((n, h, w, c, kh, kw, k, pad, s))
    return out


@pytest.mark.parametrize("policy", [32 + 256, 32 + 128], ids=["im2col", "halo"])
def test_two_sm_kernels_on_random_geometries(policy):
    """Fuzz the 2-SM TMA kernels (im2col walk and halo-reuse strips) against the float64 oracle:
    40 random geometries per variant -- image sizes that are not multiples of anything, windows
    wider than the image under SAME padding, single-pixel images, strides (which send dgrad /
    the halo variant back to the other kernels: the dispatch is part of what is tested)."""
    from oracle import numpy_path as orc

    old = sp.get_precision()
    sp.set_precision("tf32")
    core.debug_set_kernel_policy(policy)
    try:
        for i, (n, h, w, c, kh, kw, k, pad, s) in enumerate(_random_geometries(2024, 40)):
            rng = np.random.default_rng(100 + i)
            x, wt = rng.random((n, h, w, c)) - 0.5, rng.random((kh, kw, c, k)) - 0.5
            xv, wv = sp.Variable(x), sp.Variable(wt)
            y = sp.conv2d(xv, wv, pad, s)
            gy = rng.random(y.shape) - 0.5
            g = sp.get_gradients(sp.mul(y, sp.Variable(gy)))
            y0, dx0, dw0 = orc.conv2d_grads(x, wt, gy, pad, s)
            what = f"geometry {i}: x{(n, h, w, c)} w{(kh, kw, c, k)} {pad} s{s}"
            assert_close(y.array, y0, RTOL["tf32"], what + " y")
            assert_close(g[xv], dx0, RTOL["tf32"], what + " dX")
            assert_close(g[wv], dw0, RTOL["tf32"], what + " dW")
    finally:
        core.debug_set_kernel_policy(0)
        sp.set_precision(old)


def test_rgb_first_layer_kernels_on_random_geometries():
    """Fuzz the first-layer (C = 3, 3x3) kernels in TF32 mode -- the tensor-core filter gradient
    (mma.sync over padded-linear rows) and the CUDA-core forward -- against the float64 oracle:
    image sizes from 1x1 up, both paddings, every eligible K, batches that leave tiles ragged.
    Stride-2 and K = 8 cases exercise the dispatch back to the CUDA-core gradient kernel."""
    from oracle import numpy_path as orc

    old = sp.get_precision()
    sp.set_precision("tf32")
    core.debug_set_kernel_policy(0)
    rng0 = np.random.default_rng(77)
    try:
        for i in range(36):
            n = int(rng0.choice([1, 2, 7, 33]))
            h, w = int(rng0.integers(1, 41)), int(rng0.integers(1, 41))
            k = int(rng0.choice([8, 16, 32, 64]))
            pad = str(rng0.choice(["SAME", "VALID"]))
            s = (2, 2) if rng0.random() < 0.15 else (1, 1)
            if pad == "VALID" and (h < 3 or w < 3):
                continue
            rng = np.random.default_rng(500 + i)
            x, wt = rng.random((n, h, w, 3)) - 0.5, rng.random((3, 3, 3, k)) - 0.5
            xv, wv = sp.Variable(x), sp.Variable(wt)
            y = sp.conv2d(xv, wv, pad, s)
            gy = rng.random(y.shape) - 0.5
            g = sp.get_gradients(sp.mul(y, sp.Variable(gy)))
            y0, dx0, dw0 = orc.conv2d_grads(x, wt, gy, pad, s)
            what = f"rgb geometry {i}: x{(n, h, w, 3)} k{k} {pad} s{s}"
            assert_close(y.array, y0, RTOL["tf32"], what + " y")
            assert_close(g[wv], dw0, RTOL["tf32"], what + " dW")
            assert_close(g[xv], dx0, RTOL["tf32"], what + " dX")
            # deterministic: fixed-order reductions
            g2 = sp.get_gradients(sp.mul(sp.conv2d(xv, wv, pad, s), sp.Variable(gy)))
            assert np.array_equal(np.asarray(g[wv]), np.asarray(g2[wv])), what
    finally:
        sp.set_precision(old)
