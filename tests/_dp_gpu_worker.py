"""Worker for tests/test_gpu_multi.py: one rank of an N-GPU NCCL job.  Trains the CIFAR CNN
data-parallel on its shard of a global batch and writes its final weights; rank 0 also runs
the same global batch on a single GPU for comparison."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import smallpebble_b200 as sp  # noqa: E402
from smallpebble_b200 import workloads  # noqa: E402


def main():
    out_dir = sys.argv[1]
    sp.set_precision("fp32")
    sp.distributed.init()
    rank, world = sp.distributed.rank(), sp.distributed.world_size()
    # step 1: single all-reduce, 2: split planned, 3+: overlapped (eager); with the steady-state
    # replay on, steps 4-5 record the CUDA graphs (NCCL all-reduce inside) and 6+ replay them
    steps, per_gpu = int(os.environ.get("SP_TEST_STEPS", "6")), 16
    wl = workloads.cifar_cnn(seed=0)
    adam = sp.Adam()
    losses = []
    for step in range(steps):
        x, y = workloads.synthetic_batch("cnn", per_gpu * world, seed=100 + step)
        xs, ys = sp.distributed.shard(x, y)
        loss_val, _ = wl.train_step(adam, xs, ys)
        losses.append(sp.distributed.allreduce_scalar(float(loss_val.array)))
    overlapped = sp.distributed._overlap["used"]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), losses=np.array(losses),
             overlapped_steps=np.array(overlapped),
             replayed_updates=np.array(sp.steady.stats["replays_u"]),
             peer_exchange=np.array(int(sp.distributed._peer["handle"] is not None)),
             **{f"w{i}": np.asarray(p.array) for i, p in enumerate(wl.learnables)})
    sp.distributed.barrier()
    sp.distributed.shutdown()
    if rank == 0:
        # single-GPU run on the full global batch, same seeds (eager path)
        sp.set_steady_replay(False)
        wl1 = workloads.cifar_cnn(seed=0)
        adam1 = sp.Adam()
        losses1 = []
        for step in range(steps):
            x, y = workloads.synthetic_batch("cnn", per_gpu * world, seed=100 + step)
            loss_val, _ = wl1.train_step(adam1, x, y)
            losses1.append(float(loss_val.array))
        np.savez(os.path.join(out_dir, "single.npz"), losses=np.array(losses1),
                 **{f"w{i}": np.asarray(p.array) for i, p in enumerate(wl1.learnables)})
    print(f"rank {rank}/{world} ok")


if __name__ == "__main__":
    main()
