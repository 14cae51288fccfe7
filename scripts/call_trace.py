"""Names of the kernel-launching ABI calls of one eager training step, per API call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import smallpebble_b200 as sp
from smallpebble_b200 import workloads, steady, _abi
from smallpebble_b200.array import DeviceArray
kind = sys.argv[1] if len(sys.argv) > 1 else "cnn"
batch = {"cnn": 128, "mlp": 200, "vgg": 64}[kind]
steady.set_enabled(False)
x, y = workloads.synthetic_batch(kind, batch)
wl = workloads.BUILDERS[kind](seed=0)
adam = sp.Adam()
xd = DeviceArray.from_host(x); yd = DeviceArray.from_host(y, dtype=np.int32); xd.ptr; yd.ptr
def step(trace):
    wl.x_in.assign_value(sp.Variable(xd)); wl.y_true.assign_value(yd)
    n0 = len(_abi.prof.records); lv = wl.loss.run(); lv.array.ptr
    n1 = len(_abi.prof.records); g = sp.get_gradients(lv)
    n2 = len(_abi.prof.records); adam.training_step(wl.learnables, g)
    n3 = len(_abi.prof.records)
    if trace:
        torch.cuda.synchronize()
        for name, a, b in (("run", n0, n1), ("get_gradients", n1, n2), ("training_step", n2, n3)):
            print(name)
            for r in _abi.prof.records[a:b]:
                print(f"   {r[0]:32s} {r[3].elapsed_time(r[4]) * 1e3:7.1f} us")
for i in range(4): step(False)
_abi.prof.active = True
step(False); _abi.prof.records.clear()
step(True)
