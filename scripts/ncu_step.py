"""One eager CIFAR/MLP/VGG training step repeated (for ncu launch lists): SP_STEADY=0."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SP_STEADY", "0")
import torch
import smallpebble_b200 as sp
from smallpebble_b200 import workloads
from smallpebble_b200.array import DeviceArray
kind = sys.argv[1] if len(sys.argv) > 1 else "cnn"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else {"cnn": 128, "mlp": 200, "vgg": 256}[kind]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
if len(sys.argv) > 4:
    sp.set_precision(sys.argv[4])
x, y = workloads.synthetic_batch(kind, batch)
wl = workloads.BUILDERS[kind](seed=0)
adam = sp.Adam()
xd = DeviceArray.from_host(x); yd = DeviceArray.from_host(y, dtype=np.int32); xd.ptr; yd.ptr
for i in range(steps):
    if i == steps - 1:  # under `ncu --profile-from-start off`: only the last step is captured
        torch.cuda.synchronize(); torch.cuda.profiler.start()
        DeviceArray.full((4,), 0.0)  # (the first launch after cudaProfilerStart is not captured)
    wl.x_in.assign_value(sp.Variable(xd)); wl.y_true.assign_value(yd)
    lv = wl.loss.run(); adam.training_step(wl.learnables, sp.get_gradients(lv))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(lv.array))
