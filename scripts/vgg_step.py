"""A few VGG-6 (64x64, batch 256) training steps: the workload of BASELINE configs[3], for the
ncu launch list (per-kernel share of the step)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallpebble_b200 as sp
from smallpebble_b200 import workloads
from smallpebble_b200.array import DeviceArray

world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:  # torchrun: data parallel, 256 images per GPU (weak scaling)
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.dup2(2, 1)  # NCCL prints its banner on stdout
    sp.distributed.init()
x, y = workloads.synthetic_batch("vgg", 256)
wl = workloads.vgg6(seed=0)
adam = sp.Adam()
xd, yd = DeviceArray.from_host(x), DeviceArray.from_host(y, dtype=np.int32)
xd.ptr, yd.ptr
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(steps):
    if i == steps // 2:
        e0.record()
    wl.x_in.assign_value(sp.Variable(xd)); wl.y_true.assign_value(yd)
    lv = wl.loss.run(); adam.training_step(wl.learnables, sp.get_gradients(lv))
e1.record()
torch.cuda.synchronize()
msg = (f"loss {float(lv.array)} ms/step over the last {steps - steps // 2} steps: "
       f"{e0.elapsed_time(e1) / (steps - steps // 2)}")
if world > 1:
    msg += f" | rank {sp.distributed.rank()}/{world} overlapped exchanges {sp.distributed._overlap['used']}"
    sys.stderr.write(msg + "\n")
    sp.distributed.barrier()
    torch.cuda.synchronize()
    os._exit(0)
print(msg)
