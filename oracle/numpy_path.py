"""ORACLE (test infrastructure, NOT product code) -- float64 NumPy restatement of
SmallPebble's training hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this module; the product package `smallpebble_b200`
never does (tests/test_no_oracle_in_product.py enforces it).

Parity status: PINNED.  `oracle/pin_against_reference.py` imports the unmodified
reference from /root/reference in the build container and checks every function
here against it (forward values, every gradient, argmax indices, index maps, Adam
steps); `tests/golden/make_golden.py` stores reference outputs as fixtures under
tests/golden/ that `tests/test_oracle_golden.py` replays on any machine.

The reference's arithmetic lives in NumPy (un-vendored third party; declared
`numpy>=1.24.0`, pyproject.toml:18-20, locked 2.2.6/2.4.1 in uv.lock). This file
restates the *algorithm the reference builds from NumPy calls*, citing the
reference line each piece follows ("sp.py" = smallpebble/smallpebble.py):

  tape node + per-path reverse sweep ........ sp.py:36-87
  broadcast bookkeeping / unbroadcast-sum ... sp.py:177-193, 659-684
  conv2d = pad -> window view -> im2col -> GEMM  sp.py:124-163
  window view + np.add.at col2im ............ sp.py:378-390, 687-718
  TF-"SAME" padding amounts ................. sp.py:721-761
  maxpool2d / argmax reduce ................. sp.py:257-300
  matmul, leaky_relu, softmax, cross_entropy  sp.py:239-247, 231-236, 357-364, 612-621
  Adam / SGD ................................ sp.py:552-583, 645-653
  initialisers .............................. sp.py:596-635

The structure is deliberately different from the reference (a Node class with
explicit vjp lists, direct-form helper functions used by the parity tests) but the
numerical recipe -- operation order, float64, NumPy primitives used -- is the same,
so results agree to the last few ulps and CPU cost is representative.
"""

from __future__ import annotations

import math
from collections import defaultdict

import numpy as np

# --------------------------------------------------------------------------- #
# integer helpers (bit-exact contract)
# --------------------------------------------------------------------------- #


def same_pad_1d(length: int, stride: int, kern: int) -> tuple[int, int]:
    """TF 'SAME' padding for one axis (sp.py:721-730): extra pixel goes to the end."""
    n_out = -(-length // stride)
    total = max((n_out - 1) * stride + kern - length, 0)
    front = total // 2
    return front, total - front


def window_out_size(length: int, stride: int, kern: int) -> int:
    """Number of window positions on an already padded axis (sp.py:712-714)."""
    return math.ceil((length - kern + 1) / stride)


def conv_geometry(h, w, kh, kw, sy, sx, padding):
    """(pad_top, pad_bottom, pad_left, pad_right, out_h, out_w) for conv2d/maxpool2d
    (sp.py:733-761 + 712-714)."""
    if padding == "SAME":
        pt, pb = same_pad_1d(h, sy, kh)
        pl, pr = same_pad_1d(w, sx, kw)
    elif padding == "VALID":
        pt = pb = pl = pr = 0
    else:
        raise ValueError("padding must be 'SAME' or 'VALID'.")
    oh = window_out_size(h + pt + pb, sy, kh)
    ow = window_out_size(w + pl + pr, sx, kw)
    return pt, pb, pl, pr, oh, ow


def im2col_index_map(imh, imw, kh, kw, sy, sx):
    """Row/col index of every patch element of one 2-D image (sp.py:764-782).
    Returns ((rows, cols), out_h, out_w, n_patches); rows/cols are int64
    [n_patches, kh*kw]."""
    ys = np.arange(0, imh - kh + 1, sy, dtype=np.int64)
    xs = np.arange(0, imw - kw + 1, sx, dtype=np.int64)
    dy, dx = np.divmod(np.arange(kh * kw, dtype=np.int64), kw)
    rows = (ys[:, None, None] + dy[None, None, :]) + 0 * xs[None, :, None]
    cols = (xs[None, :, None] + dx[None, None, :]) + 0 * ys[:, None, None]
    oh, ow = len(ys), len(xs)
    return (rows.reshape(oh * ow, kh * kw), cols.reshape(oh * ow, kh * kw)), oh, ow, oh * ow


def broadcast_reduce_axes(shape_a, shape_b):
    """Axes of the broadcast result over which each operand was added/repeated
    (sp.py:659-684)."""
    nd = max(len(shape_a), len(shape_b))
    ea = (1,) * (nd - len(shape_a)) + tuple(shape_a)
    eb = (1,) * (nd - len(shape_b)) + tuple(shape_b)
    for p, q in zip(ea, eb):
        if not (p == q or p == 1 or q == 1):
            raise ValueError(f"could not broadcast shapes {shape_a} {shape_b}")
    ra = tuple(i for i in range(nd) if i < nd - len(shape_a) or (ea[i] == 1 and eb[i] > 1))
    rb = tuple(i for i in range(nd) if i < nd - len(shape_b) or (eb[i] == 1 and ea[i] > 1))
    return ra, rb


def window_view(x: np.ndarray, window, strides) -> np.ndarray:
    """Zero-copy strided window view, shape = reduced + window (sp.py:687-718)."""
    if len(window) != x.ndim or len(strides) != x.ndim:
        raise ValueError("Must provide one window size and one stride for each dimension of x.")
    if any(v < 0 for v in window) or any(v < 0 for v in strides):
        raise ValueError("window/strides cannot contain negative values")
    if any(n < k for n, k in zip(x.shape, window)):
        raise ValueError("window shape cannot be larger than input array shape")
    reduced = tuple(math.ceil((n - k + 1) / s) for n, s, k in zip(x.shape, strides, window))
    byte_steps = tuple(b * s for b, s in zip(x.strides, strides))
    return np.lib.stride_tricks.as_strided(
        x, shape=reduced + tuple(window), strides=byte_steps + x.strides
    )


# --------------------------------------------------------------------------- #
# tape (sp.py:36-87): node = value + [(parent, vjp)], reverse sweep walks every
# root->leaf PATH (no memoisation), accumulating `grads[parent] += vjp(path)`.
# --------------------------------------------------------------------------- #


class Node:
    __slots__ = ("value", "edges", "is_learnable")

    def __init__(self, value, edges=()):
        self.value = value
        self.edges = list(edges)

    @property
    def shape(self):
        return self.value.shape


def backprop(root: Node) -> dict:
    """Per-path depth-first reverse sweep, as sp.py:77-87."""
    grads = defaultdict(lambda: 0)

    def walk(node, path_value):
        for parent, vjp in node.edges:
            contribution = vjp(path_value)
            grads[parent] += contribution
            walk(parent, contribution)

    grads[root] = np.ones(root.value.shape, root.value.dtype)
    walk(root, grads[root])
    return dict(grads)


def _broadcast_wrappers(a: Node, b: Node, skip_last=0):
    """Extra nodes that sum an incoming gradient back to each operand's own shape
    (sp.py:177-193): sum over repeat axes, reshape, then `zeros + g`."""
    sa = a.value.shape[: a.value.ndim - skip_last] if skip_last else a.value.shape
    sb = b.value.shape[: b.value.ndim - skip_last] if skip_last else b.value.shape
    ra, rb = broadcast_reduce_axes(sa, sb)

    def make(node, axes):
        def vjp(g):
            g = np.sum(g, axis=axes).reshape(node.value.shape)
            return np.zeros(node.value.shape, node.value.dtype) + g

        return Node(node.value, [(node, vjp)])

    return make(a, ra), make(b, rb)


# ---- primitive ops --------------------------------------------------------- #


def op_add(a, b):  # sp.py:100-108
    out = a.value + b.value
    wa, wb = _broadcast_wrappers(a, b)
    return Node(out, [(wa, lambda g: g), (wb, lambda g: g)])


def op_sub(a, b):  # sp.py:393-401
    out = a.value - b.value
    wa, wb = _broadcast_wrappers(a, b)
    return Node(out, [(wa, lambda g: g), (wb, lambda g: -g)])


def op_mul(a, b):  # sp.py:303-311
    out = a.value * b.value
    wa, wb = _broadcast_wrappers(a, b)
    return Node(out, [(wa, lambda g: g * b.value), (wb, lambda g: g * a.value)])


def op_div(a, b):  # sp.py:166-174
    out = a.value / b.value
    wa, wb = _broadcast_wrappers(a, b)
    return Node(
        out,
        [(wa, lambda g: g / b.value), (wb, lambda g: -g * a.value / np.square(b.value))],
    )


def op_neg(a):  # sp.py:314-318
    return Node(-a.value, [(a, lambda g: -g)])


def op_exp(a):  # sp.py:196-200
    return Node(np.exp(a.value), [(a, lambda g: g * np.exp(a.value))])


def op_log(a):  # sp.py:224-228
    return Node(np.log(a.value), [(a, lambda g: g / a.value)])


def op_square(a):  # sp.py:367-375
    return Node(np.square(a.value), [(a, lambda g: g * 2 * a.value)])


def op_sum(a, axis=None):  # sp.py:404-415 (incl. the `if axis:` truthiness quirk)
    def vjp(g):
        base = np.zeros(a.value.shape, a.value.dtype)
        if axis:
            g = np.expand_dims(g, axis)
        return base + g

    return Node(np.sum(a.value, axis), [(a, vjp)])


def op_reshape(a, shape):  # sp.py:333-337
    return Node(np.reshape(a.value, shape), [(a, lambda g: g.reshape(a.value.shape))])


def op_getitem(a, index):  # sp.py:210-221
    def vjp(g):
        out = np.zeros(a.value.shape, a.value.dtype)
        np.add.at(out, index, g)
        return out

    return Node(a.value[index], [(a, vjp)])


def op_pad(a, pad_width):  # sp.py:321-330
    def vjp(g):
        return g[tuple(slice(lo, g.shape[i] - hi) for i, (lo, hi) in enumerate(pad_width))]

    return Node(np.pad(a.value, pad_width), [(a, vjp)])


def op_window_view(a, window, strides):  # sp.py:378-390 (np.add.at col2im)
    def vjp(g):
        out = np.zeros(a.value.shape, a.value.dtype)
        np.add.at(window_view(out, window, strides), None, g)
        return out

    return Node(window_view(a.value, window, strides), [(a, vjp)])


def op_matmul(a, b):  # sp.py:239-247
    out = np.matmul(a.value, b.value)
    wa, wb = _broadcast_wrappers(a, b, skip_last=2)
    return Node(
        out,
        [
            (wa, lambda g: np.matmul(g, np.swapaxes(b.value, -2, -1))),
            (wb, lambda g: np.matmul(np.swapaxes(a.value, -2, -1), g)),
        ],
    )


def op_leaky_relu(a, alpha=0.02):  # sp.py:231-236
    dt = a.value.dtype
    slope = np.where(a.value > 0, np.array(1, dt), np.array(alpha, dt))
    return Node(a.value * slope, [(a, lambda g: g * slope)])


def op_max_axis(a, axis):  # sp.py:257-279 (first max wins, one-hot is float64)
    nd = a.value.ndim
    axis = axis if axis >= 0 else nd + axis
    moved = np.swapaxes(a.value, axis, -1)
    flat = moved.reshape(-1, moved.shape[-1])
    which = np.argmax(flat, axis=-1)
    best = np.take_along_axis(flat, which[:, None], -1)
    out_shape = tuple(1 if i == axis else n for i, n in enumerate(a.value.shape))

    def vjp(g):
        onehot = np.zeros(flat.shape)
        onehot[np.arange(flat.shape[0]), which] = 1
        moved_shape = list(a.value.shape)
        moved_shape[axis], moved_shape[-1] = moved_shape[-1], moved_shape[axis]
        return g * np.swapaxes(onehot.reshape(moved_shape), axis, -1)

    node = Node(best.reshape(out_shape), [(a, vjp)])
    return node, which


# ---- composite ops (the reference builds these from the primitives) -------- #


def _pad_images(images, padding, kh, kw, sy, sx):  # sp.py:733-761
    _, h, w, _ = images.value.shape
    pt, pb, pl, pr, _, _ = conv_geometry(h, w, kh, kw, sy, sx, padding)
    if padding == "SAME":
        images = op_pad(images, [(0, 0), (pt, pb), (pl, pr), (0, 0)])
    return images


def op_conv2d(images, kernels, padding="SAME", strides=(1, 1)):  # sp.py:124-163
    n = images.value.shape[0]
    kh, kw, cin, cout = kernels.value.shape
    sy, sx = strides
    images = _pad_images(images, padding, kh, kw, sy, sx)
    patches = op_window_view(images, (1, kh, kw, cin), (1, sy, sx, 1))
    oh, ow = patches.value.shape[1], patches.value.shape[2]
    lhs = op_reshape(patches, [n * oh * ow, kh * kw * cin])
    rhs = op_reshape(kernels, [kh * kw * cin, cout])
    return op_reshape(op_matmul(lhs, rhs), [n, oh, ow, cout])


def op_maxpool2d(images, kh, kw, padding="SAME", strides=(1, 1), return_index=False):
    # sp.py:282-300
    sy, sx = strides
    images = _pad_images(images, padding, kh, kw, sy, sx)
    patches = op_window_view(images, (1, kh, kw, 1), [1, sy, sx, 1])
    flat = op_reshape(patches, patches.value.shape[:4] + (-1,))
    best, which = op_max_axis(flat, axis=-1)
    out = op_reshape(best, best.value.shape[:-1])
    return (out, which.reshape(out.value.shape)) if return_index else out


def op_softmax(a, axis=-1):  # sp.py:357-364 -- shift by the max of the WHOLE array
    shifted = op_exp(op_sub(a, Node(np.max(a.value))))
    keep = list(a.value.shape)
    keep[axis] = 1
    return op_div(shifted, op_reshape(op_sum(shifted, axis=axis), keep))


def op_cross_entropy(probs, labels):  # sp.py:612-621 -- sum reduction
    picked = op_getitem(probs, (np.arange(len(labels)), labels))
    return op_neg(op_sum(op_log(picked)))


# --------------------------------------------------------------------------- #
# direct-form helpers used by the parity tests: plain arrays in, arrays out
# --------------------------------------------------------------------------- #


def _f64(x):
    return np.asarray(x, dtype=np.float64)


def conv2d_forward(x, w, padding="SAME", strides=(1, 1)):
    return op_conv2d(Node(_f64(x)), Node(_f64(w)), padding, strides).value


def conv2d_grads(x, w, gy, padding="SAME", strides=(1, 1)):
    """(y, dX, dW) for upstream gradient gy (the tape seeds ones, so multiply first)."""
    xn, wn = Node(_f64(x)), Node(_f64(w))
    y = op_conv2d(xn, wn, padding, strides)
    g = backprop(op_mul(y, Node(_f64(gy))))
    return y.value, g[xn], g[wn]


def maxpool2d_forward(x, kh, kw, padding="SAME", strides=(1, 1)):
    """(y, idx) with idx = dy*kw+dx of the first maximum of each (zero padded) window."""
    y, idx = op_maxpool2d(Node(_f64(x)), kh, kw, padding, strides, return_index=True)
    return y.value, idx.astype(np.int64)


def maxpool2d_grads(x, kh, kw, gy, padding="SAME", strides=(1, 1)):
    xn = Node(_f64(x))
    y = op_maxpool2d(xn, kh, kw, padding, strides)
    g = backprop(op_mul(y, Node(_f64(gy))))
    return y.value, g[xn]


def matmul_grads(a, b, gy):
    an, bn = Node(_f64(a)), Node(_f64(b))
    y = op_matmul(an, bn)
    g = backprop(op_mul(y, Node(_f64(gy))))
    return y.value, g[an], g[bn]


def leaky_relu_grads(x, gy, alpha=0.02):
    xn = Node(_f64(x))
    y = op_leaky_relu(xn, alpha)
    return y.value, backprop(op_mul(y, Node(_f64(gy))))[xn]


def softmax_grads(x, gy, axis=-1):
    xn = Node(_f64(x))
    y = op_softmax(xn, axis)
    return y.value, backprop(op_mul(y, Node(_f64(gy))))[xn]


def softmax_xent_grads(logits, labels):
    """(probs, loss, dlogits) for loss = cross_entropy(softmax(logits), labels)."""
    xn = Node(_f64(logits))
    p = op_softmax(xn)
    loss = op_cross_entropy(p, np.asarray(labels))
    return p.value, loss.value, backprop(loss)[xn]


def bias_add_grads(a, bias, gy):
    an, bn = Node(_f64(a)), Node(_f64(bias))
    y = op_add(an, bn)
    g = backprop(op_mul(y, Node(_f64(gy))))
    return y.value, g[an], g[bn]


# --------------------------------------------------------------------------- #
# optimisers (sp.py:552-583, 645-653)
# --------------------------------------------------------------------------- #


class AdamState:
    def __init__(self, alpha=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps
        self.t = defaultdict(int)
        self.m = defaultdict(lambda: 0)
        self.v = defaultdict(lambda: 0)

    def step(self, nodes, grads):
        for p in nodes:
            self.t[p] += 1
            g = grads[p]
            self.m[p] = self.beta1 * self.m[p] + (1 - self.beta1) * g
            self.v[p] = self.beta2 * self.v[p] + (1 - self.beta2) * g**2
            m_hat = self.m[p] / (1 - self.beta1 ** self.t[p])
            v_hat = self.v[p] / (1 - self.beta2 ** self.t[p])
            p.value = p.value - self.alpha * m_hat / (np.sqrt(v_hat) + self.eps)


def sgd_update(nodes, grads, lr=1e-3):
    for p in nodes:
        p.value = p.value - lr * grads[p]


# --------------------------------------------------------------------------- #
# initialisers (global NumPy RNG, call order matters: sp.py:596-635)
# --------------------------------------------------------------------------- #


def conv_kernel_init(kh, kw, cin, cout):  # sp.py:605-606
    sigma = np.sqrt(6 / (kh * kw * cin + kh * kw * cout))
    return sigma * (np.random.random([kh, kw, cin, cout]) - 0.5)


def linear_init(n_in, n_out):  # sp.py:624-627, 632-633
    sigma = np.sqrt(4 / (n_in + n_out))
    w = np.random.random([n_in, n_out]) * sigma - sigma / 2
    return w, np.ones([n_out], np.float32)


# --------------------------------------------------------------------------- #
# the BASELINE.json models, written against the tape above
# --------------------------------------------------------------------------- #


class _Model:
    params: list

    def logits(self, x: Node) -> Node:
        raise NotImplementedError

    def loss(self, x, labels):
        return op_cross_entropy(op_softmax(self.logits(Node(_f64(x)))), np.asarray(labels))

    def predict(self, x):
        return op_softmax(self.logits(Node(_f64(x)))).value

    def train_step(self, opt: AdamState, x, labels):
        """One reference-style step: forward, per-path backward, Adam.  Returns loss."""
        loss = self.loss(x, labels)
        grads = backprop(loss)
        opt.step(self.params, grads)
        return float(loss.value), grads


class MnistMlp(_Model):
    """README.md:122-137: 784-100-100-10, leaky_relu, softmax."""

    def __init__(self, sizes=(784, 100, 100, 10)):
        self.layers = []
        for n_in, n_out in zip(sizes[:-1], sizes[1:]):
            w, b = linear_init(n_in, n_out)
            self.layers.append((Node(w), Node(b)))
        self.params = [p for wb in self.layers for p in wb]

    def logits(self, h):
        for i, (w, b) in enumerate(self.layers):
            h = op_add(op_matmul(h, w), b)
            if i + 1 < len(self.layers):
                h = op_leaky_relu(h)
        return h


class CifarCnn(_Model):
    """README.md:261-281: 3x[conv3x3 VALID + leaky_relu + maxpool 2x2/2 SAME] + linear."""

    def __init__(self, channels=(3, 32, 128, 128), n_classes=10, flat=3 * 3 * 128):
        self.kernels = [
            Node(conv_kernel_init(3, 3, cin, cout)) for cin, cout in zip(channels[:-1], channels[1:])
        ]
        w, b = linear_init(flat, n_classes)
        self.fc = (Node(w), Node(b))
        self.flat = flat
        self.params = self.kernels + list(self.fc)

    def logits(self, h):
        for k in self.kernels:
            h = op_conv2d(h, k, "VALID", (1, 1))
            h = op_leaky_relu(h)
            h = op_maxpool2d(h, 2, 2, "SAME", (2, 2))
        h = op_reshape(h, [-1, self.flat])
        return op_add(op_matmul(h, self.fc[0]), self.fc[1])


class Vgg6(_Model):
    """BASELINE.json config 4: 64-64-p-128-128-p-256-256-p-linear, 3x3 SAME, on 64x64x3."""

    def __init__(self, n_classes=10, image=64, widths=(64, 64, 128, 128, 256, 256)):
        self.kernels, cin = [], 3
        for cout in widths:
            self.kernels.append(Node(conv_kernel_init(3, 3, cin, cout)))
            cin = cout
        self.flat = (image // 8) ** 2 * widths[-1]
        w, b = linear_init(self.flat, n_classes)
        self.fc = (Node(w), Node(b))
        self.params = self.kernels + list(self.fc)

    def logits(self, h):
        for i, k in enumerate(self.kernels):
            h = op_leaky_relu(op_conv2d(h, k, "SAME", (1, 1)))
            if i % 2 == 1:
                h = op_maxpool2d(h, 2, 2, "SAME", (2, 2))
        h = op_reshape(h, [-1, self.flat])
        return op_add(op_matmul(h, self.fc[0]), self.fc[1])


def synthetic_batch(kind: str, batch: int, seed: int = 0, image: int = 64):
    """`rng = default_rng(seed)`; X uniform [0,1), labels 0..9."""
    rng = np.random.default_rng(seed)
    if kind == "mlp":
        x = rng.random([batch, 784])
    elif kind == "cnn":
        x = rng.random([batch, 32, 32, 3])
    elif kind == "vgg":
        x = rng.random([batch, image, image, 3])
    else:
        raise ValueError(kind)
    return x, rng.integers(0, 10, batch)


def build_model(kind: str, seed: int = 0, **kw):
    np.random.seed(seed)
    return {"mlp": MnistMlp, "cnn": CifarCnn, "vgg": Vgg6}[kind](**kw)
