"""Step time of the unchanged loop with steady-state replay on/off (device-timed)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallpebble_b200 as sp
from smallpebble_b200 import workloads, steady
from smallpebble_b200.array import DeviceArray

kind = sys.argv[1] if len(sys.argv) > 1 else "cnn"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else {"cnn": 128, "mlp": 200, "vgg": 256}[kind]
if len(sys.argv) > 3:
    sp.set_precision(sys.argv[3])
x, y = workloads.synthetic_batch(kind, batch)
wl = workloads.BUILDERS[kind](seed=0)
adam = sp.Adam()
nb = 8
xs = [DeviceArray.from_host(np.random.default_rng(i).random(x.shape)) for i in range(nb)]
yd = DeviceArray.from_host(y, dtype=np.int32)
[a.ptr for a in xs], yd.ptr
pin = torch.empty((nb,) + x.shape, dtype=torch.float64).pin_memory().numpy()
for i in range(nb):
    pin[i] = np.random.default_rng(i).random(x.shape)

def step(i):
    wl.x_in.assign_value(sp.Variable(xs[i % nb])); wl.y_true.assign_value(yd)
    lv = wl.loss.run(); adam.training_step(wl.learnables, sp.get_gradients(lv))

def step_e2e(i):
    wl.x_in.assign_value(sp.Variable(pin[i % nb])); wl.y_true.assign_value(y)
    lv = wl.loss.run()
    if np.isnan(lv.array): raise RuntimeError
    adam.training_step(wl.learnables, sp.get_gradients(lv))

n = 300 if kind != "vgg" else 30
for fn, name in ((step, "resident"), (step_e2e, "e2e")):
    for i in range(30): fn(i)
    torch.cuda.synchronize()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for i in range(n): fn(i)
        host = time.perf_counter() - t0
        e1.record(); torch.cuda.synchronize()
        print(f"{kind} {name}: {e0.elapsed_time(e1)/n*1e3:.1f} us/step device-timed, host issue {host/n*1e6:.1f} us/step  | {batch*n/(e0.elapsed_time(e1)/1e3):.0f} img/s")
print("steady", steady.enabled, "fork", steady.fork_leaf_grads, steady.stats)
