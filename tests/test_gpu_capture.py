"""GPU: a step replayed from a CUDA graph (sp.capture) must train exactly like the eager
step -- same kernels, same order, same inputs => bitwise the same weights."""

import numpy as np
import pytest

import smallpebble_b200 as sp
from smallpebble_b200 import workloads

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,batch", [("mlp", 64), ("cnn", 16)])
def test_captured_step_equals_eager(kind, batch):
    rng = np.random.default_rng(3)
    batches = []
    for _ in range(6):
        x, y = workloads.synthetic_batch(kind, batch, seed=int(rng.integers(1 << 30)))
        batches.append((x, y))

    # eager reference: batch 0 twice (the recording pass = 1 warm-up + 1 capture run, neither
    # of which... see below) -- simplest exact protocol: eager runs b0, b0, b0, b2..b5
    wl_e = workloads.BUILDERS[kind](seed=0)
    adam_e = sp.Adam()
    seq = [batches[0]] * 2 + batches[2:]
    eager_losses = [float(wl_e.train_step(adam_e, x, y)[0].array) for x, y in seq]

    wl_c = workloads.BUILDERS[kind](seed=0)
    adam_c = sp.Adam()
    step = wl_c.captured_train_step(adam_c)
    step.warmup = 1
    # first call: 1 eager warm-up on b0 (applied), 1 capture pass (recorded, NOT executed),
    # then the replay on b0 (applied) => two updates with b0, like `seq`
    cap_losses = [float(step(*batches[0]).array)]
    for x, y in batches[2:]:
        cap_losses.append(float(step(x, y).array))
    assert np.allclose(cap_losses, eager_losses[1:], rtol=0, atol=0), (cap_losses, eager_losses)
    for pe, pc in zip(wl_e.learnables, wl_c.learnables):
        assert np.array_equal(np.asarray(pe.array), np.asarray(pc.array))
    assert adam_c.t[wl_c.learnables[0]] == adam_e.t[wl_e.learnables[0]] == len(seq)
    assert step.kernels_per_replay > 10
