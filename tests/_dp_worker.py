"""Worker for tests/test_distributed_gloo.py: one rank of a world_size-N gloo job on CPU.
Checks the data-parallel plumbing (shard -> per-rank gradient -> flat bucket -> all-reduce)
against the oracle's global-batch gradient."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import cases as C  # noqa: E402
import smallpebble_b200 as sp  # noqa: E402
from oracle import numpy_path as orc  # noqa: E402
from smallpebble_b200 import distributed as dist  # noqa: E402
from smallpebble_b200.array import DeviceArray  # noqa: E402


def main():
    dist.init(backend="gloo")
    rank, world = dist.rank(), dist.world_size()
    x, y = C.synthetic_batch("mlp", 50)
    # every rank builds the same model from the same seed (weights replicated)
    model = orc.build_model("mlp", seed=0)
    full_loss = model.loss(x, y)
    full_grads = orc.backprop(full_loss)
    xs, ys = dist.shard(x, y)
    lo, hi = dist.shard_bounds(len(x))
    assert len(xs) == hi - lo
    shard_loss = model.loss(xs, ys)
    g = orc.backprop(shard_loss)
    local = [DeviceArray.from_host(g[p]) for p in model.params]  # host-pending arrays
    summed = dist.allreduce_gradients(local)
    for p, s in zip(model.params, summed):
        ref = np.asarray(full_grads[p], dtype=np.float64)
        got = np.asarray(s, dtype=np.float64)
        assert got.shape == ref.shape
        err = np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30)
        assert err < 1e-5, f"rank {rank}: all-reduced gradient differs from global batch: {err}"
    total = dist.allreduce_scalar(float(shard_loss.value))
    assert abs(total - float(full_loss.value)) < 1e-8 * abs(float(full_loss.value))
    # parameter broadcast: rank 1 perturbs, rank 0's values win
    v = sp.Variable(np.full((3, 2), float(rank)))
    dist.broadcast_parameters([v], src=0)
    assert np.all(np.asarray(v.array) == 0.0)
    dist.barrier()
    dist.shutdown()
    print(f"rank {rank}/{world} ok")


if __name__ == "__main__":
    main()
