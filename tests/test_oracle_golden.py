"""CPU: the oracle (oracle/numpy_path.py) must reproduce the golden vectors that
tests/golden/make_golden.py recorded from the unmodified reference."""

import numpy as np

import cases as C
from oracle import numpy_path as orc

TOL = 1e-12


def close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b) / (1 + np.abs(b)), initial=0.0) <= tol


def test_conv2d(golden):
    for case in C.CONV_CASES:
        name, xs, ws, strides, pad = case
        x, w, gy = C.conv_inputs(case)
        y, dx, dw = orc.conv2d_grads(x, w, gy, pad, strides)
        close(y, golden[f"conv/{name}/y"])
        close(dx, golden[f"conv/{name}/dx"])
        close(dw, golden[f"conv/{name}/dw"])


def test_maxpool2d(golden):
    for case in C.POOL_CASES:
        name, xs, (kh, kw), strides, pad = case
        x, gy = C.pool_inputs(case)
        y, dx = orc.maxpool2d_grads(x, kh, kw, gy, pad, strides)
        y2, idx = orc.maxpool2d_forward(x, kh, kw, pad, strides)
        close(y, golden[f"pool/{name}/y"], 0)
        close(dx, golden[f"pool/{name}/dx"], 0)
        assert np.array_equal(idx, golden[f"pool/{name}/idx"].astype(np.int64))


def test_matmul(golden):
    for case in C.MATMUL_CASES:
        a, b, gy = C.matmul_inputs(case)
        y, da, db = orc.matmul_grads(a, b, gy)
        close(y, golden[f"matmul/{case[0]}/y"])
        close(da, golden[f"matmul/{case[0]}/da"])
        close(db, golden[f"matmul/{case[0]}/db"])


def test_softmax_xent_leaky_bias(golden):
    for case in C.SOFTMAX_CASES:
        x, gy = C.softmax_inputs(case)
        y, dx = orc.softmax_grads(x, gy, case[2])
        close(y, golden[f"softmax/{case[0]}/y"])
        close(dx, golden[f"softmax/{case[0]}/dx"])
    logits, labels = C.xent_inputs()
    p, loss, dl = orc.softmax_xent_grads(logits, labels)
    close(p, golden["xent/probs"])
    close(loss, golden["xent/loss"])
    close(dl, golden["xent/dlogits"])
    x, gy = C.leaky_inputs()
    for alpha in (0.02, 0.1):
        y, dx = orc.leaky_relu_grads(x, gy, alpha)
        close(y, golden[f"leaky/{alpha}/y"], 0)
        close(dx, golden[f"leaky/{alpha}/dx"], 0)
    a, bias, gy = C.bias_inputs()
    y, da, db = orc.bias_add_grads(a, bias, gy)
    close(y, golden["bias/y"])
    close(da, golden["bias/da"])
    close(db, golden["bias/db"])


def test_adam_sgd(golden):
    p0, grads = C.adam_inputs()
    node = orc.Node(p0.copy())
    opt = orc.AdamState()
    for i, g in enumerate(grads):
        opt.step([node], {node: g})
        close(node.value, golden[f"adam/p{i + 1}"])
    # SURVEY A8 known answer: constant gradient => |step| ~= alpha
    node = orc.Node(np.array([1.0, -2.0, 3.0]))
    opt = orc.AdamState()
    g = np.array([0.1, -0.2, 0.3])
    opt.step([node], {node: g})
    close(node.value, [0.999, -1.999, 2.999], 1e-7)
    opt.step([node], {node: g})
    close(node.value, [0.998, -1.998, 2.998], 1e-7)
    node = orc.Node(p0.copy())
    orc.sgd_update([node], {node: grads[0]}, lr=0.01)
    close(node.value, golden["sgd/p1"])


def test_integer_helpers(golden):
    pads = [orc.same_pad_1d(L, s, k) for L in range(1, 40) for s in (1, 2, 3, 4, 7, 10)
            for k in (1, 2, 3, 5, 13, 22)]
    assert np.array_equal(np.array(pads, np.int64), golden["int/pad_amounts"])
    for args in [(5, 6, 2, 3, 2, 1), (32, 32, 3, 3, 1, 1), (13, 17, 5, 2, 3, 4)]:
        (r, c), oh, ow, n = orc.im2col_index_map(*args)
        key = "int/patches_index/" + "_".join(map(str, args))
        assert np.array_equal(r, golden[key + "/rows"]) and r.dtype == np.int64
        assert np.array_equal(c, golden[key + "/cols"])
        assert [oh, ow, n] == list(golden[key + "/meta"])
    # SURVEY A2 examples
    assert orc.same_pad_1d(32, 1, 3) == (1, 1) and orc.same_pad_1d(13, 2, 2) == (0, 1)
    assert orc.same_pad_1d(32, 2, 5) == (1, 2) and orc.same_pad_1d(7, 4, 2) == (0, 0)


def test_models(golden):
    for kind, batch, steps in C.MODEL_CASES:
        x, y = C.synthetic_batch(kind, batch)
        model = orc.build_model(kind, seed=0)
        opt = orc.AdamState()
        losses = []
        for step in range(steps):
            loss, grads = model.train_step(opt, x, y)
            if step == 0:
                for i, p in enumerate(model.params):
                    close(np.asarray(grads[p]).ravel()[::7], golden[f"model/{kind}/grad0/{i}"], 1e-10)
            losses.append(loss)
        close(losses, golden[f"model/{kind}/losses"], 1e-11)
        for i, p in enumerate(model.params):
            close(p.value.ravel()[::7], golden[f"model/{kind}/w_final/{i}"], 1e-10)
        close(model.predict(x), golden[f"model/{kind}/pred_final"], 1e-10)


def test_maxpool_zero_padding_semantics():
    """all-negative window touching the SAME pad returns 0.0 and routes its
    gradient into the pad (dropped); all-equal window -> index 0; NaN wins."""
    x = -np.ones((1, 3, 3, 1))
    y, idx = orc.maxpool2d_forward(x, 2, 2, "SAME", (2, 2))
    assert y[0, 1, 1, 0] == 0.0 and y[0, 0, 0, 0] == -1.0 and idx[0, 0, 0, 0] == 0
    _, dx = orc.maxpool2d_grads(x, 2, 2, np.ones((1, 2, 2, 1)), "SAME", (2, 2))
    assert dx.sum() == 1.0  # only the window with no padding keeps its gradient
    x = np.zeros((1, 2, 2, 1))
    x[0, 1, 0, 0] = np.nan
    y, idx = orc.maxpool2d_forward(x, 2, 2, "VALID", (2, 2))
    assert np.isnan(y[0, 0, 0, 0]) and idx[0, 0, 0, 0] == 2
