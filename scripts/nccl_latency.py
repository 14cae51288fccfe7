"""Latency of the one collective of the data-parallel path: in-place fp32 sum all-reduce of the
flat gradient bucket (196,714 floats for the README CNN) through smallpebble_b200.distributed,
timed on the device (CUDA events around 200 back-to-back calls, max over ranks).  Launch with
torchrun; NCCL_ALGO / NCCL_PROTO in the environment select the variant under test."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from smallpebble_b200 import distributed as spd  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
real_stdout = os.dup(1)
os.dup2(2, 1)  # NCCL prints its banner on stdout
spd.init()
for count in (196_716, 1_000_000):
    buf = torch.ones(count, dtype=torch.float32, device="cuda")
    for _ in range(20):
        spd._allreduce_sum_f32(buf.data_ptr(), count, buf)
        buf.fill_(1.0)
    torch.cuda.synchronize()
    dist.barrier()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            spd._allreduce_sum_f32(buf.data_ptr(), count, buf)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 200)
        buf.fill_(1.0)
    t = torch.tensor([best], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if spd.rank() == 0:
        os.write(real_stdout, (f"world {spd.world_size()} floats {count} ALGO={os.environ.get('NCCL_ALGO', '-')} "
                               f"PROTO={os.environ.get('NCCL_PROTO', '-')}: {1e3 * float(t.item()):.1f} us per all-reduce\n").encode())
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
