"""GPU: the reference's own test-suite (tests/test_smallpebble.py, 35 tests) re-run against
the device implementation through the unchanged API.  Same inputs, same structure; the
absolute 1e-6 float64 tolerance becomes the relative fp32/TF32 tolerance of BASELINE.json,
check_grad is the fixed one (gradients really are compared), and the three TensorFlow
comparisons use the oracle (pinned against the reference; torch-verified in SURVEY.md)."""

import numpy as np
import pytest

import smallpebble_b200 as sp
from conftest import RTOL, assert_close
from oracle import numpy_path as orc
from smallpebble_b200.gradcheck import check_grad, numgrads, rel_rmse

pytestmark = pytest.mark.gpu


def test_add():
    check_grad([np.random.random([20, 20]), np.random.random([20, 20])], sp.add, np.add)


def test_add_broadcast():
    check_grad([np.random.random([4, 2, 3]), np.random.random([3, 4, 1, 2, 3])], sp.add, np.add)


def test_add_broadcast_scalar():
    check_grad([np.array(4.0), np.random.random([3, 4, 1, 2, 3])], sp.add, np.add)


def test_add_operator():
    func = lambda a, b: a + b  # noqa: E731
    check_grad([np.random.random([20, 20]), np.random.random([20, 20])], func, func)


def test_add_at():
    len_a, len_b = 5, 10
    args = [np.random.random([len_a]), np.random.random([len_b])]
    indices = np.random.randint(0, len_a, size=[len_b])

    def np_func(a, b):
        result = a.copy()
        np.add.at(result, indices, b)
        return result

    check_grad(args, lambda a, b: sp.add_at(a, indices, b), np_func)


def test_div_broadcast_operator():
    args = [np.random.random([3, 2, 5]) * 1000, np.random.random([4, 3, 1, 5]) * 2 + 1]
    func = lambda a, b: a / b  # noqa: E731
    check_grad(args, func, func, delta=1e-5)


def test_exponential():
    check_grad([np.random.random([3, 2, 5])], sp.exp, np.exp, delta=1e-5)


def test_expand_dims():
    axis = [3]
    check_grad([np.random.random([3, 2, 5])], lambda a: sp.expand_dims(a, axis),
               lambda a: np.expand_dims(a, axis))


def test_getitem():
    indices = (slice(None), slice(1), (0, 2))
    check_grad([np.random.random([3, 2, 5])], lambda a: sp.getitem(a, indices),
               lambda a: a[indices])


def test_getitem_dunder():
    func = lambda a: a[0, :, 0:2]  # noqa: E731
    check_grad([np.random.random([4, 5, 3])], func, func)


def test_log():
    check_grad([np.random.random([3, 2, 5]) * 100 + 1], sp.log, np.log, delta=1e-5)


def test_leaky_relu():
    alpha = 0.1
    check_grad([np.random.random([8, 8, 8]) * 10 - 5], lambda a: sp.leaky_relu(a, alpha),
               lambda a: np.maximum(a, a * alpha), delta=1e-5)


def test_matmul():
    check_grad([np.random.random([2, 5]), np.random.random([5, 9])], sp.matmul, np.matmul)


def test_matmul_broadcast():
    check_grad([np.random.random([2, 4, 3, 2, 5]), np.random.random([1, 3, 5, 9])],
               sp.matmul, np.matmul)


def test_matrix_transpose():
    check_grad([np.random.random([3, 2, 5])], sp.matrix_transpose,
               lambda a: np.swapaxes(a, -2, -1))


def test_maxax():
    axis = -2
    check_grad([np.random.random([3, 2, 5, 4])], lambda a: sp.maxax(a, axis),
               lambda a: np.max(a, axis, keepdims=True), delta=1e-5)


def test_mul_broadcast_operator():
    func = lambda a, b: a * b  # noqa: E731
    check_grad([np.random.random([4, 1, 3, 2, 5]), np.random.random([5, 3, 1, 5])], func, func)


def test_neg_operator():
    func = lambda a: -a  # noqa: E731
    check_grad([np.random.random([5, 3, 1, 5])], func, func)


def test_pad():
    pad_width = ((2, 1), (4, 4), (0, 1))
    check_grad([np.random.random([3, 2, 5])], lambda a: sp.pad(a, pad_width),
               lambda a: np.pad(a, pad_width))


def test_reshape():
    shape = (6, 5, 1)
    check_grad([np.random.random([3, 2, 5])], lambda a: sp.reshape(a, shape),
               lambda a: a.reshape(shape))


def test_setat():
    args = [np.random.random([3, 2, 5]), np.array(5)]
    indices = (slice(None), slice(1), (0, 2))

    def np_func(a, b):
        result = a.copy()
        result[indices] = b
        return result

    check_grad(args, lambda a, b: sp.setat(a, indices, b), np_func)


def test_softmax():
    np_func = lambda a: np.exp(a) / np.sum(np.exp(a), axis=-1, keepdims=True)  # noqa: E731
    check_grad([np.random.random([100, 10])], sp.softmax, np_func, delta=1e-6, eps=1e-6)


def test_square():
    check_grad([np.random.random([3, 2, 5])], sp.square, np.square, delta=1e-5)


def test_sub_broadcast_operator():
    func = lambda a, b: a - b  # noqa: E731
    check_grad([np.random.random([4, 2, 3]), np.random.random([4, 1, 2, 3])], func, func)


def test_sum():
    axis = (1, 3)
    check_grad([np.random.random([4, 3, 2, 5])], lambda a: sp.sum(a, axis),
               lambda a: np.sum(a, axis))


def test_sum_axis_variants():
    x = np.random.random([4, 3, 5])
    for axis in (None, 0, (0,), 1, -1, (0, 1), (0, 2)):
        check_grad([x], lambda a: sp.sum(a, axis), lambda a: np.sum(a, axis))


def test_where():
    args = [np.random.random([4, 2, 3]), np.random.random([3, 4, 1, 2, 3])]
    condition = args[0] > args[1]
    check_grad(args, lambda a, b: sp.where(condition, a, b),
               lambda a, b: np.where(condition, a, b))


def test_strided_sliding_view():
    x = np.random.random([2, 7, 6, 3])
    window, strides = (1, 3, 2, 1), (1, 2, 1, 1)
    check_grad([x], lambda a: sp.strided_sliding_view(a, window, strides),
               lambda a: sp.np_strided_sliding_view(a, window, strides).copy())


# ---------------- HIGHER OPS (reference: vs TensorFlow; here: vs the pinned oracle)


def generate_images_and_kernels(imagedims, kerndims):
    np.random.seed(0)
    n_images, imheight, imwidth, _ = imagedims
    kernheight, kernwidth, channels_in, channels_out = kerndims
    images = np.random.random([n_images, imheight, imwidth, channels_in])
    kernels = np.random.random([kernheight, kernwidth, channels_in, channels_out])
    return images, kernels


def test_conv2d_result(precision):  # ref tests:255-270
    images, kernels = generate_images_and_kernels([3, 55, 66, 5], [13, 22, 5, 7])
    for stride in range(1, 11, 3):
        for padding in ["SAME", "VALID"]:
            result = sp.conv2d(sp.Variable(images), sp.Variable(kernels), padding=padding,
                               strides=[stride, stride])
            expect = orc.conv2d_forward(images, kernels, padding, (stride, stride))
            assert_close(result.array, expect, RTOL[precision], f"s={stride} {padding}")


def test_conv2d_grads(precision):  # ref tests:274-307
    images, kernels = generate_images_and_kernels([5, 16, 12, 2], [2, 4, 2, 3])
    images_sp, kernels_sp = sp.Variable(images), sp.Variable(kernels)
    for stride in range(1, 11, 3):
        for padding in ["SAME", "VALID"]:
            result = sp.conv2d(images_sp, kernels_sp, padding=padding, strides=[stride, stride])
            g = sp.get_gradients(result)
            ones = np.ones(result.shape)
            _, dx, dw = orc.conv2d_grads(images, kernels, ones, padding, (stride, stride))
            assert_close(g[images_sp], dx, RTOL[precision], f"dX s={stride} {padding}")
            assert_close(g[kernels_sp], dw, RTOL[precision], f"dW s={stride} {padding}")


def test_maxpool2d():  # ref tests:311-338
    np.random.seed(0)
    images = np.random.random([4, 12, 14, 3])
    images_sp = sp.Variable(images)
    for stride in range(1, 11, 3):
        for padding in ["SAME", "VALID"]:
            result = sp.maxpool2d(images_sp, 2, 2, padding, [stride, stride])
            g = sp.get_gradients(result)
            ones = np.ones(result.shape)
            y, dx = orc.maxpool2d_grads(images, 2, 2, ones, padding, (stride, stride))
            assert_close(result.array, y, 1e-6, "pool y")
            assert_close(g[images_sp], dx, 1e-6, "pool dX")


# ---------------- DELAYED GRAPH EXECUTION (ref tests:351-395)


def test_operation_and_placeholder_result():
    a_array, b_array = np.ones([10, 5]), np.ones([5, 10])
    a, b = sp.Placeholder(), sp.Placeholder()
    c = sp.Lazy(sp.matmul)(a, b)
    d = sp.Lazy(sp.mul)(c, sp.Variable(np.array(5)))
    a.assign_value(sp.Variable(a_array))
    b.assign_value(sp.Variable(b_array))
    result = d.run()
    assert np.all(np.asarray(result.array - np.matmul(a_array, b_array) * 5) == 0)


def test_operation_and_placeholder_gradients():
    a_array, b_array = np.ones([10, 5]), np.ones([5, 10])
    a_sp, b_sp = sp.Variable(a_array), sp.Variable(b_array)
    a, b = sp.Placeholder(), sp.Placeholder()
    y = sp.Lazy(sp.matmul)(a, b)
    a.assign_value(a_sp)
    b.assign_value(b_sp)
    result = y.run()
    grads = sp.get_gradients(result)
    num = numgrads(np.matmul, [a_array, b_array], n=1, delta=1)
    for spval, numval in zip([result.array, grads[a_sp], grads[b_sp]],
                             [np.matmul(a_array, b_array), num[0], num[1]]):
        assert rel_rmse(spval, numval) < 1e-5


# ---------------- OPTIMISERS (ref tests:422-454)


def test_adam():
    a = sp.Variable(np.random.random([3, 5]))
    b = sp.Variable(np.random.random([5, 3]))
    y_true = sp.Variable(np.ones([3, 3]))
    adam = sp.Adam()
    losses = []
    for _ in range(100):
        loss = sp.sum(sp.square(y_true - sp.matmul(a, b)))
        losses.append(float(loss.array))
        adam.training_step([a, b], sp.get_gradients(loss))
    assert np.all(np.diff(losses) < 0), "Not converging."


def test_sgd_step():
    a = sp.Variable(np.random.random([3, 5]))
    b = sp.Variable(np.random.random([5, 3]))
    y_true = sp.Variable(np.ones([3, 3]))
    losses = []
    for _ in range(100):
        loss = sp.sum(sp.square(y_true - sp.matmul(a, b)))
        losses.append(float(loss.array))
        sp.sgd_step([a, b], sp.get_gradients(loss))
    assert np.all(np.diff(losses) < 0), "Not converging."


# ---------------- TAPE CONTRACT


def test_user_defined_op_with_numpy_closure():
    """README 'custom operations': closures over arrays, written with NumPy, still work."""
    def cube(a):
        value = np.asarray(a.array) ** 3
        return sp.Variable(value, [(a, lambda g: g * 3 * np.asarray(a.array) ** 2)])

    x = np.random.random([4, 3])
    v = sp.Variable(x)
    y = sp.sum(cube(v) * v)  # v used twice: both paths accumulate
    g = sp.get_gradients(y)
    assert_close(g[v], 4 * x**3, 1e-5, "custom op gradient")
    assert y in g and np.asarray(g[y]).shape == ()


def test_gradients_of_non_scalar_root_and_shared_nodes():
    b = sp.Variable(np.random.random([2, 3]))
    y = b * b
    g = sp.get_gradients(y)
    assert_close(g[b], 2 * np.asarray(b.array), 1e-6, "d(sum b*b)/db")
    assert np.all(np.asarray(g[y]) == 1.0)
