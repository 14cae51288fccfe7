"""The BASELINE.json workloads written against the public API exactly as the reference's
README writes them (README.md:122-137 MNIST MLP, README.md:261-281 CIFAR CNN) plus the wider
VGG-style config.  Shared by bench.py, __graft_entry__.smoke() and the model parity tests so
that all three exercise the same user-level code."""

from __future__ import annotations

import numpy as np

import smallpebble_b200 as sp


class Workload:
    """A lazy graph with its placeholders and learnables."""

    def __init__(self, x_in, y_true, y_pred, loss):
        self.x_in, self.y_true, self.y_pred, self.loss = x_in, y_true, y_pred, loss
        self.learnables = sp.get_learnables(y_pred)

    def train_step(self, optimiser, xbatch, ybatch):
        """One iteration of the README training loop (README.md:310-325) minus the eval pass."""
        self.x_in.assign_value(sp.Variable(xbatch))
        self.y_true.assign_value(ybatch)
        loss_val = self.loss.run()
        gradients = sp.get_gradients(loss_val)
        optimiser.training_step(self.learnables, gradients)
        return loss_val, gradients

    def captured_train_step(self, optimiser):
        """The same step as `train_step`, recorded once into a CUDA graph (sp.capture):
        `step(xbatch, ybatch) -> loss Variable`."""
        def step(x, y):
            self.x_in.assign_value(sp.Variable(x))
            self.y_true.assign_value(y)
            loss_val = self.loss.run()
            optimiser.training_step(self.learnables, sp.get_gradients(loss_val))
            return loss_val

        return sp.capture(step)

    def predict(self, xbatch):
        self.x_in.assign_value(sp.Variable(xbatch))
        return self.y_pred.run()


def mnist_mlp(seed: int | None = 0) -> Workload:
    if seed is not None:
        np.random.seed(seed)
    x_in, y_true = sp.Placeholder(), sp.Placeholder()
    h = sp.linearlayer(28 * 28, 100)(x_in)
    h = sp.Lazy(sp.leaky_relu)(h)
    h = sp.linearlayer(100, 100)(h)
    h = sp.Lazy(sp.leaky_relu)(h)
    h = sp.linearlayer(100, 10)(h)
    y_pred = sp.Lazy(sp.softmax)(h)
    loss = sp.Lazy(sp.cross_entropy)(y_pred, y_true)
    return Workload(x_in, y_true, y_pred, loss)


def cifar_cnn(seed: int | None = 0) -> Workload:
    if seed is not None:
        np.random.seed(seed)
    x_in, y_true = sp.Placeholder(), sp.Placeholder()
    h = sp.convlayer(height=3, width=3, depth=3, n_kernels=32)(x_in)
    h = sp.Lazy(sp.leaky_relu)(h)
    h = sp.Lazy(lambda a: sp.maxpool2d(a, 2, 2, strides=[2, 2]))(h)
    h = sp.convlayer(3, 3, 32, 128, padding="VALID")(h)
    h = sp.Lazy(sp.leaky_relu)(h)
    h = sp.Lazy(lambda a: sp.maxpool2d(a, 2, 2, strides=[2, 2]))(h)
    h = sp.convlayer(3, 3, 128, 128, padding="VALID")(h)
    h = sp.Lazy(sp.leaky_relu)(h)
    h = sp.Lazy(lambda a: sp.maxpool2d(a, 2, 2, strides=[2, 2]))(h)
    h = sp.Lazy(lambda x: sp.reshape(x, [-1, 3 * 3 * 128]))(h)
    h = sp.linearlayer(3 * 3 * 128, 10)(h)
    y_pred = sp.Lazy(sp.softmax)(h)
    loss = sp.Lazy(sp.cross_entropy)(y_pred, y_true)
    return Workload(x_in, y_true, y_pred, loss)


def vgg6(seed: int | None = 0, image: int = 64, widths=(64, 64, 128, 128, 256, 256)) -> Workload:
    """BASELINE.json config 4: 3x3 SAME convs 64-64-pool-128-128-pool-256-256-pool-linear."""
    if seed is not None:
        np.random.seed(seed)
    x_in, y_true = sp.Placeholder(), sp.Placeholder()
    h, cin = x_in, 3
    for i, cout in enumerate(widths):
        h = sp.convlayer(3, 3, cin, cout, padding="SAME")(h)
        h = sp.Lazy(sp.leaky_relu)(h)
        if i % 2 == 1:
            h = sp.Lazy(lambda a: sp.maxpool2d(a, 2, 2, strides=[2, 2]))(h)
        cin = cout
    flat = (image // 8) ** 2 * widths[-1]
    h = sp.Lazy(lambda x: sp.reshape(x, [-1, flat]))(h)
    h = sp.linearlayer(flat, 10)(h)
    y_pred = sp.Lazy(sp.softmax)(h)
    loss = sp.Lazy(sp.cross_entropy)(y_pred, y_true)
    return Workload(x_in, y_true, y_pred, loss)


def synthetic_batch(kind: str, batch: int, seed: int = 0, image: int = 64):
    """rng = default_rng(seed); X uniform [0,1) float64, labels 0..9."""
    rng = np.random.default_rng(seed)
    shape = {"mlp": [batch, 784], "cnn": [batch, 32, 32, 3], "vgg": [batch, image, image, 3]}[kind]
    x = rng.random(shape)
    return x, rng.integers(0, 10, batch)


BUILDERS = {"mlp": mnist_mlp, "cnn": cifar_cnn, "vgg": vgg6}
