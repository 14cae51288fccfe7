"""Shared case table for the golden fixtures (inputs are regenerated from seeds, only
reference OUTPUTS are stored).  Used by tests/golden/make_golden.py (which runs the real
reference in the build container) and by the parity tests (oracle on CPU, CUDA on GPU).
"""

from __future__ import annotations

import numpy as np

# name, x shape, w shape, (sy, sx), padding
CONV_CASES = [
    ("tf_result_s4_same", (3, 55, 66, 5), (13, 22, 5, 7), (4, 4), "SAME"),  # ref tests:257
    ("tf_result_s7_valid", (3, 55, 66, 5), (13, 22, 5, 7), (7, 7), "VALID"),
    ("tf_result_s10_same", (3, 55, 66, 5), (13, 22, 5, 7), (10, 10), "SAME"),
    ("tf_grads_s1_same", (5, 16, 12, 2), (2, 4, 2, 3), (1, 1), "SAME"),  # ref tests:276
    ("tf_grads_s1_valid", (5, 16, 12, 2), (2, 4, 2, 3), (1, 1), "VALID"),
    ("tf_grads_s4_same", (5, 16, 12, 2), (2, 4, 2, 3), (4, 4), "SAME"),
    ("tf_grads_s7_valid", (5, 16, 12, 2), (2, 4, 2, 3), (7, 7), "VALID"),
    ("k3_c8_same", (2, 12, 11, 8), (3, 3, 8, 16), (1, 1), "SAME"),
    ("k3_c8_s2_same", (2, 12, 11, 8), (3, 3, 8, 16), (2, 2), "SAME"),
    ("k3_c3_valid", (4, 10, 10, 3), (3, 3, 3, 32), (1, 1), "VALID"),  # CNN conv1 shape family
    ("k5_c4_s2_same", (2, 9, 13, 4), (5, 5, 4, 8), (2, 2), "SAME"),
    ("k1_c16_s2", (2, 8, 8, 16), (1, 1, 16, 8), (2, 2), "SAME"),
    ("asym_stride", (1, 9, 10, 2), (3, 2, 2, 5), (2, 1), "SAME"),
    ("asym_stride_valid", (1, 9, 10, 2), (3, 2, 2, 5), (2, 3), "VALID"),
    ("c32_k128", (2, 15, 15, 32), (3, 3, 32, 128), (1, 1), "VALID"),  # CNN conv2 shape family
    # channel counts % 32 == 0: eligible for the 2-SM TMA kernels (csrc/conv_tma2.cu)
    ("c64_k128_same", (3, 12, 11, 64), (3, 3, 64, 128), (1, 1), "SAME"),
    ("c32_k64_k5_s2", (2, 17, 19, 32), (5, 5, 32, 64), (2, 2), "SAME"),
    ("c128_k256_valid", (2, 9, 9, 128), (3, 3, 128, 256), (1, 1), "VALID"),
    ("c32_k192_k2x3", (3, 9, 10, 32), (2, 3, 32, 192), (1, 1), "SAME"),
    ("c512_k64_k1", (1, 10, 13, 512), (1, 1, 512, 64), (1, 1), "SAME"),
    # strided convs with channel counts % 32 == 0: the data gradient runs as one stride-1
    # sub-problem per phase of the input pixels (conv_tma2.cu)
    ("c32_k64_k3_s2_valid", (3, 13, 12, 32), (3, 3, 32, 64), (2, 2), "VALID"),
    ("c64_k64_k1_s2_same", (2, 9, 8, 64), (1, 1, 64, 64), (2, 2), "SAME"),  # pixels without taps
    ("c32_k64_k4_s3_same", (2, 14, 11, 32), (4, 4, 32, 64), (3, 3), "SAME"),
    ("c32_k64_k3x2_s2x1_valid", (2, 10, 9, 32), (3, 2, 32, 64), (2, 1), "VALID"),
    # RGB 3x3 first layers: register-tiled CUDA-core kernels (csrc/conv_smallc.cu)
    ("k3_c3_same_k64", (2, 9, 7, 3), (3, 3, 3, 64), (1, 1), "SAME"),
    ("k3_c3_s2_k16", (2, 11, 10, 3), (3, 3, 3, 16), (2, 2), "SAME"),
    ("k3_c3_valid_k8", (3, 6, 13, 3), (3, 3, 3, 8), (1, 1), "VALID"),
]

# name, x shape, (kh, kw), (sy, sx), padding
POOL_CASES = [
    ("tf_s1_same", (4, 12, 14, 3), (2, 2), (1, 1), "SAME"),  # ref tests:313
    ("tf_s1_valid", (4, 12, 14, 3), (2, 2), (1, 1), "VALID"),
    ("tf_s4_same", (4, 12, 14, 3), (2, 2), (4, 4), "SAME"),
    ("tf_s7_valid", (4, 12, 14, 3), (2, 2), (7, 7), "VALID"),
    ("tf_s10_same", (4, 12, 14, 3), (2, 2), (10, 10), "SAME"),
    ("cnn_pool_odd", (3, 13, 13, 8), (2, 2), (2, 2), "SAME"),  # pad 13->14, zero pad can win
    ("cnn_pool_5", (3, 5, 5, 16), (2, 2), (2, 2), "SAME"),
    ("k3_s2", (2, 9, 7, 4), (3, 3), (2, 2), "SAME"),
    ("k3_s1_overlap", (2, 9, 7, 4), (3, 3), (1, 1), "SAME"),
    ("k32_s1_valid", (2, 6, 6, 5), (3, 2), (1, 1), "VALID"),
    ("k3_s3_same", (2, 10, 8, 8), (3, 3), (3, 3), "SAME"),  # tiled backward, pad_t = 1
    ("k2_s2_valid_odd", (2, 7, 9, 4), (2, 2), (2, 2), "VALID"),  # last row/col in no window
]

# a shape, b shape
MATMUL_CASES = [
    ("plain", (2, 5), (5, 9)),  # ref tests:138
    ("broadcast", (2, 4, 3, 2, 5), (1, 3, 5, 9)),  # ref tests:143
    ("mlp_fc1", (200, 784), (784, 100)),
    ("cnn_fc", (128, 1152), (1152, 10)),
    ("odd", (37, 53), (53, 19)),
    # whole 32 / 64-column blocks: eligible for the 2-SM TMA kernels (a GEMM is a 1 x 1 conv)
    ("tma_small", (300, 96), (96, 128)),
]

# name, shape, axis (sp.py:257-279: for axis != last the reference lays the result out in
# swapped-row order without swapping back -- the golden vectors pin that, values and gradient)
MAXAX_CASES = [("last", (4, 6), -1), ("axis0_3d", (3, 4, 5), 0), ("axis1_3d", (3, 4, 5), 1),
               ("axis0_2d", (5, 7), 0)]

SOFTMAX_CASES = [("rows", (100, 10), -1), ("cols", (7, 5), 0), ("mid", (3, 4, 6), 1)]

MODEL_CASES = [("mlp", 200, 5), ("cnn", 16, 3)]  # kind, batch, Adam steps


def _rng(name: str):
    return np.random.default_rng(abs(hash_name(name)) % (2**32))


def hash_name(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def conv_inputs(case):
    name, xs, ws, strides, pad = case
    r = _rng("conv/" + name)
    x, w = r.random(xs) - 0.5, r.random(ws) - 0.5
    pt, pb, pl, pr, oh, ow = _geom(xs[1], xs[2], ws[0], ws[1], strides, pad)
    gy = r.random((xs[0], oh, ow, ws[3])) - 0.5
    return x, w, gy


def pool_inputs(case):
    name, xs, k, strides, pad = case
    r = _rng("pool/" + name)
    x = r.random(xs) - 0.5  # negatives so SAME zero padding can win the max
    x[0, :2, :2, :] = 0.25  # exact ties: first (dy,dx) wins
    pt, pb, pl, pr, oh, ow = _geom(xs[1], xs[2], k[0], k[1], strides, pad)
    gy = r.random((xs[0], oh, ow, xs[3])) - 0.5
    return x, gy


def matmul_inputs(case):
    name, sa, sb = case
    r = _rng("matmul/" + name)
    a, b = r.random(sa) - 0.5, r.random(sb) - 0.5
    gy = r.random(np.broadcast_shapes(sa[:-2], sb[:-2]) + (sa[-2], sb[-1])) - 0.5
    return a, b, gy


def maxax_inputs(case):
    name, shape, axis = case
    r = _rng("maxax/" + name)
    x = r.random(shape) - 0.5
    out_shape = tuple(1 if i == (axis % len(shape)) else v for i, v in enumerate(shape))
    return x, r.random(out_shape) - 0.5


def softmax_inputs(case):
    name, shape, axis = case
    r = _rng("softmax/" + name)
    return (r.random(shape) - 0.5) * 6, r.random(shape) - 0.5


def xent_inputs():
    r = _rng("xent")
    return (r.random((128, 10)) - 0.5) * 8, r.integers(0, 10, 128)


def leaky_inputs():
    r = _rng("leaky")
    x = r.random((8, 8, 8)) * 10 - 5
    x[0, 0, :3] = 0.0  # x == 0 takes the alpha branch
    return x, r.random(x.shape) - 0.5


def bias_inputs():
    r = _rng("bias")
    return r.random((200, 100)) - 0.5, r.random(100).astype(np.float32), r.random((200, 100)) - 0.5


def adam_inputs():
    r = _rng("adam")
    p0 = r.random((33, 7)) - 0.5
    grads = [r.random((33, 7)) - 0.5 for _ in range(4)]
    return p0, grads


def synthetic_batch(kind: str, batch: int, seed: int = 0, image: int = 64):
    """SURVEY.md 8(d) synthetic data: X uniform [0,1) float64, labels 0..9."""
    rng = np.random.default_rng(seed)
    shape = {"mlp": [batch, 784], "cnn": [batch, 32, 32, 3], "vgg": [batch, image, image, 3]}[kind]
    x = rng.random(shape)
    return x, rng.integers(0, 10, batch)


def _geom(h, w, kh, kw, strides, pad):
    import math

    def same(L, s, k):
        n = -(-L // s)
        tot = max((n - 1) * s + k - L, 0)
        return tot // 2, tot - tot // 2

    sy, sx = strides
    pt, pb = same(h, sy, kh) if pad == "SAME" else (0, 0)
    pl, pr = same(w, sx, kw) if pad == "SAME" else (0, 0)
    oh = math.ceil((h + pt + pb - kh + 1) / sy)
    ow = math.ceil((w + pl + pr - kw + 1) / sx)
    return pt, pb, pl, pr, oh, ow
