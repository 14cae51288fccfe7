"""Pin oracle/numpy_path.py against the UNMODIFIED reference imported from
/root/reference (only possible in the build container; the GPU box has no
/root/reference).  Run:  python oracle/pin_against_reference.py

Checks forward values, every gradient, integer index maps / argmax indices
(bit-exact) and multi-step Adam training of the README models.  Exit code 0 and a
final "PINNED" line mean the oracle reproduces the reference; the same inputs and
reference outputs are frozen as fixtures by tests/golden/make_golden.py.
"""

from __future__ import annotations

import os
import sys

import numpy as np

REF = os.environ.get("SP_REFERENCE_PATH", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import smallpebble as ref  # noqa: E402  (the reference)

from oracle import numpy_path as orc  # noqa: E402

TOL = 1e-12
worst = 0.0


def close(name, a, b, tol=TOL):
    global worst
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    err = float(np.max(np.abs(a - b) / (1.0 + np.abs(b)))) if a.size else 0.0
    worst = max(worst, err)
    assert err <= tol, f"{name}: max rel err {err}"
    print(f"  ok {name:55s} {err:.2e}")


def exact(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and np.array_equal(a, b), name
    print(f"  ok {name:55s} exact")


def ref_grads(fn, *arrays, gy=None):
    vs = [ref.Variable(np.array(a)) for a in arrays]
    y = fn(*vs)
    root = y if gy is None else ref.mul(y, ref.Variable(gy))
    g = ref.get_gradients(root)
    return y.array, [g[v] for v in vs]


def main():
    rng = np.random.default_rng(1234)

    print("integer helpers")
    for L in range(1, 40):
        for s in (1, 2, 3, 4, 7, 10):
            for k in (1, 2, 3, 5, 13, 22):
                assert orc.same_pad_1d(L, s, k) == ref.pad_amounts(L, s, k)
    print("  ok same_pad_1d == pad_amounts over 39*6*6 cases")
    for args in [(5, 6, 2, 3, 2, 1), (32, 32, 3, 3, 1, 1), (13, 17, 5, 2, 3, 4), (7, 7, 7, 7, 1, 1)]:
        (r0, c0), oh0, ow0, n0 = ref.patches_index(*args)
        (r1, c1), oh1, ow1, n1 = orc.im2col_index_map(*args)
        exact(f"patches_index{args} rows", r1, r0)
        exact(f"patches_index{args} cols", c1, c0)
        assert (oh0, ow0, n0) == (oh1, ow1, n1) and r1.dtype == r0.dtype
    for sa, sb in [((4, 2, 3), (3, 4, 1, 2, 3)), ((), (3, 1)), ((5, 1), (1, 5)), ((2, 4, 3), (1, 3))]:
        assert orc.broadcast_reduce_axes(sa, sb) == ref.broadcastinfo(sa, sb)
    print("  ok broadcast_reduce_axes == broadcastinfo")

    print("conv2d")
    cases = [
        ((3, 55, 66, 5), (13, 22, 5, 7), [1, 4, 7, 10]),  # tests/test_smallpebble.py:257
        ((5, 16, 12, 2), (2, 4, 2, 3), [1, 4, 7, 10]),  # tests/test_smallpebble.py:276
        ((2, 9, 10, 4), (3, 3, 4, 6), [1, 2]),
        ((2, 8, 8, 3), (5, 5, 3, 4), [1, 2]),
        ((2, 7, 6, 8), (1, 1, 8, 5), [1, 2]),
    ]
    for xs, ws, strides in cases:
        x, w = rng.random(xs) - 0.5, rng.random(ws) - 0.5
        for s in strides:
            for sy, sx in {(s, s), (s, 1), (1, s)}:
                for pad in ("SAME", "VALID"):
                    f = lambda a, b: ref.conv2d(a, b, pad, (sy, sx))  # noqa: E731
                    y0 = f(ref.Variable(x), ref.Variable(w)).array
                    gy = rng.random(y0.shape) - 0.5
                    y0, (dx0, dw0) = ref_grads(f, x, w, gy=gy)
                    y1, dx1, dw1 = orc.conv2d_grads(x, w, gy, pad, (sy, sx))
                    tag = f"{xs}x{ws} s=({sy},{sx}) {pad}"
                    close(f"conv y  {tag}", y1, y0)
                    close(f"conv dX {tag}", dx1, dx0)
                    close(f"conv dW {tag}", dw1, dw0)

    print("maxpool2d")
    for xs, (kh, kw), strides in [
        ((4, 12, 14, 3), (2, 2), [1, 4, 7, 10]),  # tests/test_smallpebble.py:313
        ((2, 13, 13, 8), (2, 2), [2]),
        ((2, 9, 7, 4), (3, 3), [1, 2, 3]),
        ((2, 6, 6, 4), (3, 2), [1, 2]),
    ]:
        x = rng.random(xs) - 0.5  # negatives: zero padding can win (sp.py:323, 265)
        x[0, :2, :2, :] = 0.25  # ties: first max wins
        for s in strides:
            for pad in ("SAME", "VALID"):
                f = lambda a: ref.maxpool2d(a, kh, kw, pad, [s, s])  # noqa: E731
                y0 = f(ref.Variable(x)).array
                gy = rng.random(y0.shape) - 0.5
                y0, (dx0,) = ref_grads(f, x, gy=gy)
                y1, dx1 = orc.maxpool2d_grads(x, kh, kw, gy, pad, (s, s))
                y2, idx = orc.maxpool2d_forward(x, kh, kw, pad, (s, s))
                # reference argmax recomputed from its own padded patch view
                pt, pb, pl, pr, oh, ow = orc.conv_geometry(xs[1], xs[2], kh, kw, s, s, pad)
                xp = np.pad(x, [(0, 0), (pt, pb), (pl, pr), (0, 0)])
                view = ref.np_strided_sliding_view(xp, (1, kh, kw, 1), (1, s, s, 1))
                idx0 = np.argmax(view.reshape(view.shape[:4] + (-1,)), axis=-1)
                tag = f"{xs} k=({kh},{kw}) s={s} {pad}"
                close(f"pool y  {tag}", y1, y0, 0)
                close(f"pool y2 {tag}", y2, y0, 0)
                exact(f"pool idx {tag}", idx, idx0)
                close(f"pool dX {tag}", dx1, dx0, 0)

    print("matmul / elementwise / softmax / xent")
    for sa, sb in [((2, 5), (5, 9)), ((2, 4, 3, 2, 5), (1, 3, 5, 9)), ((200, 784), (784, 100))]:
        a, b = rng.random(sa) - 0.5, rng.random(sb) - 0.5
        gy = rng.random(np.matmul(a, b).shape) - 0.5
        y0, (da0, db0) = ref_grads(ref.matmul, a, b, gy=gy)
        y1, da1, db1 = orc.matmul_grads(a, b, gy)
        close(f"matmul y {sa}x{sb}", y1, y0)
        close(f"matmul dA {sa}x{sb}", da1, da0)
        close(f"matmul dB {sa}x{sb}", db1, db0)
    x = rng.random((8, 8, 8)) * 10 - 5
    x[0, 0, :3] = 0.0
    gy = rng.random(x.shape) - 0.5
    for alpha in (0.02, 0.1):
        y0, (dx0,) = ref_grads(lambda a: ref.leaky_relu(a, alpha), x, gy=gy)
        y1, dx1 = orc.leaky_relu_grads(x, gy, alpha)
        close(f"leaky_relu y a={alpha}", y1, y0, 0)
        close(f"leaky_relu dX a={alpha}", dx1, dx0, 0)
    a, bias = rng.random((200, 100)) - 0.5, rng.random(100).astype(np.float32)
    gy = rng.random((200, 100)) - 0.5
    y0, (da0, db0) = ref_grads(ref.add, a, bias, gy=gy)
    y1, da1, db1 = orc.bias_add_grads(a, bias, gy)
    close("bias add y", y1, y0)
    close("bias add dA", da1, da0)
    close("bias add dbias (unbroadcast-sum)", db1, db0)
    for shape, axis in [((100, 10), -1), ((7, 5), 0), ((3, 4, 6), 1)]:
        x = (rng.random(shape) - 0.5) * 6
        gy = rng.random(shape) - 0.5
        y0, (dx0,) = ref_grads(lambda a: ref.softmax(a, axis), x, gy=gy)
        y1, dx1 = orc.softmax_grads(x, gy, axis)
        close(f"softmax y {shape} axis={axis}", y1, y0)
        close(f"softmax dX {shape} axis={axis}", dx1, dx0)
    logits = (rng.random((128, 10)) - 0.5) * 8
    labels = rng.integers(0, 10, 128)
    v = ref.Variable(logits)
    p0 = ref.softmax(v)
    l0 = ref.cross_entropy(p0, labels)
    g0 = ref.get_gradients(l0)[v]
    p1, l1, g1 = orc.softmax_xent_grads(logits, labels)
    close("softmax-xent probs", p1, p0.array)
    close("softmax-xent loss", l1, l0.array)
    close("softmax-xent dlogits", g1, g0)

    print("models: k Adam steps from identical weights")
    for kind, batch, steps in [("mlp", 200, 5), ("cnn", 16, 3)]:
        x, y = orc.synthetic_batch(kind, batch)
        om = orc.build_model(kind, seed=0)
        oopt = orc.AdamState()
        np.random.seed(0)
        xin, ytrue = ref.Placeholder(), ref.Placeholder()
        if kind == "mlp":
            h = ref.linearlayer(784, 100)(xin)
            h = ref.Lazy(ref.leaky_relu)(h)
            h = ref.linearlayer(100, 100)(h)
            h = ref.Lazy(ref.leaky_relu)(h)
            h = ref.linearlayer(100, 10)(h)
        else:
            h = ref.convlayer(3, 3, 3, 32)(xin)
            h = ref.Lazy(ref.leaky_relu)(h)
            h = ref.Lazy(lambda a: ref.maxpool2d(a, 2, 2, strides=[2, 2]))(h)
            h = ref.convlayer(3, 3, 32, 128, padding="VALID")(h)
            h = ref.Lazy(ref.leaky_relu)(h)
            h = ref.Lazy(lambda a: ref.maxpool2d(a, 2, 2, strides=[2, 2]))(h)
            h = ref.convlayer(3, 3, 128, 128, padding="VALID")(h)
            h = ref.Lazy(ref.leaky_relu)(h)
            h = ref.Lazy(lambda a: ref.maxpool2d(a, 2, 2, strides=[2, 2]))(h)
            h = ref.Lazy(lambda t: ref.reshape(t, [-1, 3 * 3 * 128]))(h)
            h = ref.linearlayer(3 * 3 * 128, 10)(h)
        ypred = ref.Lazy(ref.softmax)(h)
        loss = ref.Lazy(ref.cross_entropy)(ypred, ytrue)
        learn = ref.get_learnables(ypred)
        for i, (pr, po) in enumerate(zip(_order_like(learn, om.params), om.params)):
            exact(f"{kind} init weights[{i}]", po.value, pr.array)
        radam = ref.Adam()
        for step in range(steps):
            xin.assign_value(ref.Variable(x))
            ytrue.assign_value(y)
            lv = loss.run()
            g = ref.get_gradients(lv)
            radam.training_step(learn, g)
            lo, _ = om.train_step(oopt, x, y)
            close(f"{kind} step {step} loss", lo, lv.array, 1e-11)
        for i, (pr, po) in enumerate(zip(_order_like(learn, om.params), om.params)):
            close(f"{kind} weights[{i}] after {steps} Adam steps", po.value, pr.array, 1e-10)

    print(f"PINNED: oracle == reference (worst rel err {worst:.2e})")


def _order_like(ref_params, oracle_params):
    """get_learnables (sp.py:532-543) recurses into the first argument (the previous
    layer) before it tests this layer's weights, so its list is first-layer-first:
    the same order in which the oracle models construct their parameters."""
    assert len(ref_params) == len(oracle_params)
    return list(ref_params)


if __name__ == "__main__":
    main()
