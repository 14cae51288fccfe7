"""GPU: steady-state replay (smallpebble_b200/steady.py).  The unchanged README loop, once its
steps have settled, is replayed from CUDA graphs: everything the loop can observe must be
bit-identical to the eager path."""

import numpy as np
import pytest

import smallpebble_b200 as sp
from smallpebble_b200 import steady, workloads

pytestmark = pytest.mark.gpu


@pytest.fixture
def replay_on():
    was = steady.enabled
    yield
    steady.set_enabled(was)


def _loop(kind, batch, steps, enabled, read_leaf=False, eval_every=0):
    steady.set_enabled(enabled)
    rng = np.random.default_rng(5)
    shape = {"mlp": (batch, 784), "cnn": (batch, 32, 32, 3)}[kind]
    xs = [rng.random(shape) for _ in range(4)]
    ys = [rng.integers(0, 10, batch) for _ in range(4)]
    wl = workloads.BUILDERS[kind](seed=0)
    adam = sp.Adam()
    losses, leaf, preds = [], [], []
    for i in range(steps):
        xv = sp.Variable(xs[i % 4])
        wl.x_in.assign_value(xv)
        wl.y_true.assign_value(ys[i % 4])
        loss_val = wl.loss.run()
        assert not np.isnan(loss_val.array)
        losses.append(loss_val.array)  # README: the ARRAYS are kept, read after the loop
        g = sp.get_gradients(loss_val)
        adam.training_step(wl.learnables, g)
        if read_leaf and i % 3 == 2:
            leaf.append(np.asarray(g[xv]))  # after the update: must see the step's weights
        if eval_every and i % eval_every == 0:
            wl.x_in.assign_value(sp.Variable(xs[(i + 1) % 4][: batch // 2]))
            preds.append(np.asarray(wl.y_pred.run().array))
    weights = [np.asarray(p.array) for p in wl.learnables]
    steps_done = [adam.t[p] for p in wl.learnables]
    return [float(np.asarray(a)) for a in losses], weights, leaf, preds, steps_done


@pytest.mark.parametrize("kind,batch", [("mlp", 50), ("cnn", 16)])
def test_replayed_loop_is_bit_identical_to_eager(kind, batch, replay_on):
    steps = 14
    before = dict(steady.stats)
    got = _loop(kind, batch, steps, True, read_leaf=True, eval_every=4)
    replays = steady.stats["replays_u"] - before["replays_u"]
    assert replays >= steps - steady.WARM_STEPS - 3, steady.stats
    assert steady.stats["replays_b"] - before["replays_b"] >= replays
    assert steady.stats["aborted"] == before["aborted"]
    want = _loop(kind, batch, steps, False, read_leaf=True, eval_every=4)
    assert got[0] == want[0]  # every loss, read after the loop from the kept arrays
    for a, b in zip(got[1], want[1]):
        assert np.array_equal(a, b)
    assert len(got[2]) == len(want[2]) > 0
    for a, b in zip(got[2], want[2]):
        assert np.array_equal(a, b)
    for a, b in zip(got[3], want[3]):
        assert np.array_equal(a, b)
    assert got[4] == want[4] == [steps] * len(got[4])


def test_replayed_loop_from_page_locked_batches_reads_every_loss_right(replay_on):
    """The end-to-end README loop: float64 batches in page-locked memory (their DMA starts when
    they are wrapped, array.py `_prestage`), the loss read every step (published to a polled
    host slot by the replayed forward pass, steady.py) and AGAIN long after the slot ring has
    come round -- every value equals the eager loop's, bit for bit."""
    import torch

    rng = np.random.default_rng(6)
    batch, steps = 16, 40  # (the slot ring holds 16 results per recording)
    pins = torch.empty((4, batch, 32, 32, 3), dtype=torch.float64).pin_memory().numpy()
    pins[...] = rng.random(pins.shape)
    ys = [rng.integers(0, 10, batch) for _ in range(4)]

    def loop(enabled):
        steady.set_enabled(enabled)
        wl = workloads.cifar_cnn(seed=0)
        adam = sp.Adam()
        now, kept = [], []
        for i in range(steps):
            wl.x_in.assign_value(sp.Variable(pins[i % 4]))
            wl.y_true.assign_value(ys[i % 4])
            loss_val = wl.loss.run()
            now.append(float(np.asarray(loss_val.array)))   # read at once (polls the host slot)
            kept.append(loss_val.array)                       # ... and kept for later
            adam.training_step(wl.learnables, sp.get_gradients(loss_val))
        later = [float(np.asarray(a)) for a in kept]         # most slots have been reused by now
        return now, later, [np.asarray(p.array) for p in wl.learnables]

    before = dict(steady.stats)
    got = loop(True)
    assert steady.stats["replays_f"] - before["replays_f"] >= steps - 8
    want = loop(False)
    assert got[0] == want[0] and got[1] == want[1] and got[0] == got[1]
    for a, b in zip(got[2], want[2]):
        assert np.array_equal(a, b)


def test_replay_falls_back_when_the_pattern_changes(replay_on):
    """New batch size, a rebound parameter, a second optimiser: each is served by the eager
    path (and re-recorded later); results equal the eager loop's."""

    def run(enabled):
        steady.set_enabled(enabled)
        rng = np.random.default_rng(7)
        wl = workloads.mnist_mlp(seed=0)
        adam, other = sp.Adam(), sp.Adam(alpha=2e-3)
        out = []
        for i in range(24):
            b = 32 if i < 9 or i > 15 else 48
            wl.x_in.assign_value(sp.Variable(rng.random((b, 784))))
            wl.y_true.assign_value(rng.integers(0, 10, b))
            loss_val = wl.loss.run()
            g = sp.get_gradients(loss_val)
            if i == 12:
                p = wl.learnables[0]
                p.array = np.asarray(p.array) * 0.5
                g = sp.get_gradients(wl.loss.run())
            (other if i in (6, 20) else adam).training_step(wl.learnables, g)
            out.append(float(loss_val.array))
        return out, [np.asarray(p.array) for p in wl.learnables]

    got, want = run(True), run(False)
    assert got[0] == want[0]
    for a, b in zip(got[1], want[1]):
        assert np.array_equal(a, b)


def test_stale_result_is_refused(replay_on):
    steady.set_enabled(True)
    rng = np.random.default_rng(9)
    wl = workloads.mnist_mlp(seed=0)
    adam = sp.Adam()
    old = None
    for i in range(10):
        wl.x_in.assign_value(sp.Variable(rng.random((20, 784))))
        wl.y_true.assign_value(rng.integers(0, 10, 20))
        loss_val = wl.loss.run()
        adam.training_step(wl.learnables, sp.get_gradients(loss_val))
        if i == 8:
            old = loss_val
    if getattr(old, "_steady", None) is not None:
        with pytest.raises(RuntimeError):
            sp.get_gradients(old)


def test_host_data_created_inside_the_step_keeps_the_loop_eager(replay_on):
    """A lazy function that builds a Variable from fresh host data every run (a random mask)
    must not be frozen into a graph: the recording is abandoned and every step stays eager."""
    steady.set_enabled(True)
    rng = np.random.default_rng(1)
    masks = []

    def masked(a):
        m = (rng.random((4, 6)) > 0.5).astype(np.float64)
        masks.append(m)
        return sp.mul(a, sp.Variable(m))

    x_in = sp.Placeholder()
    w = sp.learnable(sp.Variable(np.ones((4, 6))))
    out = sp.Lazy(masked)(sp.Lazy(sp.mul)(x_in, w))
    before = steady.stats["aborted"]
    for i in range(8):
        x = np.full((4, 6), float(i + 1))
        x_in.assign_value(sp.Variable(x))
        y = out.run()
        assert np.array_equal(np.asarray(y.array), (x * masks[-1]).astype(np.float32)), i
    assert len(masks) == 8 or steady.stats["aborted"] > before
    assert steady.stats["aborted"] == before + 1
