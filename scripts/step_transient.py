"""Per-step host issue time and device time of the first steps of a timed region (why a short
`bench.py --steps K` reads slower than a long one)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallpebble_b200 as sp
from smallpebble_b200 import workloads
from smallpebble_b200.array import DeviceArray

x, y = workloads.synthetic_batch("cnn", 128)
wl = workloads.cifar_cnn(seed=0)
adam = sp.Adam()
xs = [DeviceArray.from_host(np.random.default_rng(i).random(x.shape)) for i in range(16)]
yd = DeviceArray.from_host(y, dtype=np.int32)
[a.ptr for a in xs]; yd.ptr

def step(i):
    wl.x_in.assign_value(sp.Variable(xs[i % 16])); wl.y_true.assign_value(yd)
    lv = wl.loss.run(); adam.training_step(wl.learnables, sp.get_gradients(lv))

for i in range(5): step(i)
torch.cuda.synchronize()
for trial in range(2):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    host = []
    torch.cuda.synchronize()
    evs[0].record()
    for i in range(40):
        t0 = time.perf_counter(); step(i); host.append(time.perf_counter() - t0); evs[i + 1].record()
    torch.cuda.synchronize()
    dev = [evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(40)]
    print("host us:", [round(h * 1e6) for h in host[:14]], "... mean last 20:", round(np.mean(host[20:]) * 1e6))
    print("dev  us:", [round(d) for d in dev[:14]], "... mean last 20:", round(np.mean(dev[20:])))
    print("total ms", evs[0].elapsed_time(evs[40]))
