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
print("loss", float(lv.array), "ms/step over the last", steps - steps // 2, "steps:",
      e0.elapsed_time(e1) / (steps - steps // 2))
