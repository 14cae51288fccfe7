"""GPU (needs >= 2 devices, skipped otherwise): the data-parallel run over N GPUs equals the
single-GPU run at the same global batch -- sum-reduced loss => all-reduce(SUM) of shard
gradients is exactly the global-batch gradient."""

import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                    reason="needs at least 2 GPUs")
@pytest.mark.parametrize("overlap,steady,steps,peer",
                         [("1", "0", 6, "0"), ("0", "0", 6, "0"), ("0", "1", 12, "0"),
                          ("0", "0", 6, "1"), ("0", "1", 12, "1")],
                         ids=["nccl_overlapped", "nccl_single_allreduce", "nccl_steady_replay",
                              "peer_memory_eager", "peer_memory_steady_replay"])
def test_data_parallel_equals_single_gpu(tmp_path, overlap, steady, steps, peer):
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "_dp_gpu_worker.py"),
           str(tmp_path)]
    # "1" forces the two-part exchange overlapped with the backward pass (by default only
    # large gradient buckets use it), "0" the single all-reduce on the compute stream
    # steady "1": the loop is replayed from CUDA graphs that contain the NCCL all-reduce
    # peer "1": no collective library call at all -- the fused Adam kernel of each rank sums the
    # ranks' gradient buckets over NVLink peer memory (csrc/optim.cu)
    env = dict(os.environ, SP_DP_OVERLAP=overlap, SP_STEADY=steady, SP_TEST_STEPS=str(steps),
               SP_DP_PEER=peer)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    single = np.load(tmp_path / "single.npz")
    ranks = [np.load(tmp_path / f"rank{i}.npz") for i in range(world)]
    # overlapped: the early gradients were reduced on the side stream from the third step on
    assert all(int(r["overlapped_steps"]) == (4 if overlap == "1" else 0) for r in ranks)
    if steady == "1":
        assert all(int(r["replayed_updates"]) >= steps - 6 for r in ranks)
    else:
        assert all(int(r["replayed_updates"]) == 0 for r in ranks)
    assert all(int(r["peer_exchange"]) == int(peer) for r in ranks)
    for key in single.files:
        # every rank ends with identical weights ...
        assert np.array_equal(ranks[0][key], ranks[1][key]), key
        # ... equal to the single-GPU global-batch run up to fp32 summation order
        ref = single[key].astype(np.float64)
        err = np.max(np.abs(ranks[0][key] - ref)) / max(np.max(np.abs(ref)), 1e-30)
        assert err < 2e-3 if key.startswith("w") else err < 1e-5, (key, err)
