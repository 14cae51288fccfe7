"""Generate tests/golden/hotpath_golden.npz by running the UNMODIFIED reference
(imported from /root/reference -- build container only) on the seeded inputs of
tests/golden/cases.py.  The reference cannot travel to the GPU box, the fixture can.

    python tests/golden/make_golden.py
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("SP_REFERENCE_PATH", "/root/reference"))
sys.path.insert(0, HERE)

import cases as C  # noqa: E402
import smallpebble as ref  # noqa: E402  (the reference)


def grads_of(fn, arrays, gy):
    vs = [ref.Variable(np.array(a)) for a in arrays]
    y = fn(*vs)
    g = ref.get_gradients(ref.mul(y, ref.Variable(gy)))
    return y.array, [g[v] for v in vs]


def build_ref_model(kind):
    np.random.seed(0)
    xin, ytrue = ref.Placeholder(), ref.Placeholder()
    if kind == "mlp":
        h = ref.linearlayer(784, 100)(xin)
        h = ref.Lazy(ref.leaky_relu)(h)
        h = ref.linearlayer(100, 100)(h)
        h = ref.Lazy(ref.leaky_relu)(h)
        h = ref.linearlayer(100, 10)(h)
    else:
        h = ref.convlayer(height=3, width=3, depth=3, n_kernels=32)(xin)
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
    return xin, ytrue, ypred, loss, ref.get_learnables(ypred)


def main():
    out = {}
    for case in C.CONV_CASES:
        name, xs, ws, strides, pad = case
        x, w, gy = C.conv_inputs(case)
        y, (dx, dw) = grads_of(lambda a, b: ref.conv2d(a, b, pad, strides), [x, w], gy)
        out[f"conv/{name}/y"], out[f"conv/{name}/dx"], out[f"conv/{name}/dw"] = y, dx, dw
    for case in C.POOL_CASES:
        name, xs, (kh, kw), strides, pad = case
        x, gy = C.pool_inputs(case)
        y, (dx,) = grads_of(lambda a: ref.maxpool2d(a, kh, kw, pad, list(strides)), [x], gy)
        # the argmax the reference takes (sp.py:265) on its own padded window view
        pt, pb, pl, pr, oh, ow = C._geom(xs[1], xs[2], kh, kw, strides, pad)
        xp = np.pad(x, [(0, 0), (pt, pb), (pl, pr), (0, 0)])
        view = ref.np_strided_sliding_view(xp, (1, kh, kw, 1), (1, strides[0], strides[1], 1))
        idx = np.argmax(view.reshape(view.shape[:4] + (-1,)), axis=-1)
        out[f"pool/{name}/y"], out[f"pool/{name}/dx"] = y, dx
        out[f"pool/{name}/idx"] = idx.astype(np.uint8)
    for case in C.MATMUL_CASES:
        name = case[0]
        a, b, gy = C.matmul_inputs(case)
        y, (da, db) = grads_of(ref.matmul, [a, b], gy)
        out[f"matmul/{name}/y"], out[f"matmul/{name}/da"], out[f"matmul/{name}/db"] = y, da, db
    for case in C.MAXAX_CASES:
        name, shape, axis = case
        x, gy = C.maxax_inputs(case)
        y, (dx,) = grads_of(lambda a: ref.maxax(a, axis), [x], gy)
        out[f"maxax/{name}/y"], out[f"maxax/{name}/dx"] = y, dx
    for case in C.SOFTMAX_CASES:
        name, shape, axis = case
        x, gy = C.softmax_inputs(case)
        y, (dx,) = grads_of(lambda a: ref.softmax(a, axis), [x], gy)
        out[f"softmax/{name}/y"], out[f"softmax/{name}/dx"] = y, dx
    logits, labels = C.xent_inputs()
    v = ref.Variable(logits)
    p = ref.softmax(v)
    loss = ref.cross_entropy(p, labels)
    out["xent/probs"], out["xent/loss"] = p.array, loss.array
    out["xent/dlogits"] = ref.get_gradients(loss)[v]
    x, gy = C.leaky_inputs()
    for alpha in (0.02, 0.1):
        y, (dx,) = grads_of(lambda a: ref.leaky_relu(a, alpha), [x], gy)
        out[f"leaky/{alpha}/y"], out[f"leaky/{alpha}/dx"] = y, dx
    a, bias, gy = C.bias_inputs()
    y, (da, db) = grads_of(ref.add, [a, bias], gy)
    out["bias/y"], out["bias/da"], out["bias/db"] = y, da, db
    p0, grads = C.adam_inputs()
    pv = ref.Variable(p0.copy())
    adam = ref.Adam()
    for i, g in enumerate(grads):
        adam.training_step([pv], {pv: g})
        out[f"adam/p{i + 1}"] = pv.array.copy()
    pv = ref.Variable(p0.copy())
    ref.sgd_step([pv], {pv: grads[0]}, learning_rate=0.01)
    out["sgd/p1"] = pv.array
    # integer helpers
    pads = [
        ref.pad_amounts(L, s, k)
        for L in range(1, 40)
        for s in (1, 2, 3, 4, 7, 10)
        for k in (1, 2, 3, 5, 13, 22)
    ]
    out["int/pad_amounts"] = np.array(pads, dtype=np.int64)
    for args in [(5, 6, 2, 3, 2, 1), (32, 32, 3, 3, 1, 1), (13, 17, 5, 2, 3, 4)]:
        (r, c), oh, ow, n = ref.patches_index(*args)
        key = "int/patches_index/" + "_".join(map(str, args))
        out[key + "/rows"], out[key + "/cols"] = r, c
        out[key + "/meta"] = np.array([oh, ow, n], dtype=np.int64)
    # models: loss curve, first-step gradients and final weights (strided subsample)
    for kind, batch, steps in C.MODEL_CASES:
        x, y = C.synthetic_batch(kind, batch)
        xin, ytrue, ypred, loss, learn = build_ref_model(kind)
        adam = ref.Adam()
        losses = []
        for step in range(steps):
            xin.assign_value(ref.Variable(x))
            ytrue.assign_value(y)
            lv = loss.run()
            g = ref.get_gradients(lv)
            if step == 0:
                for i, p in enumerate(learn):
                    out[f"model/{kind}/grad0/{i}"] = np.asarray(g[p]).ravel()[::7].copy()
            adam.training_step(learn, g)
            losses.append(float(lv.array))
        out[f"model/{kind}/losses"] = np.array(losses)
        for i, p in enumerate(learn):
            out[f"model/{kind}/w_final/{i}"] = np.asarray(p.array).ravel()[::7].copy()
        xin.assign_value(ref.Variable(x))
        out[f"model/{kind}/pred_final"] = ypred.run().array
    path = os.path.join(HERE, "hotpath_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
