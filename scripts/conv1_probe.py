import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallpebble_b200 as sp
from smallpebble_b200.array import DeviceArray
rng = np.random.default_rng(0)
x = sp.Variable(rng.random((128, 32, 32, 3), dtype=np.float32))
w = sp.Variable(rng.random((3, 3, 3, 32), dtype=np.float32))
for _ in range(3):
    y = sp.conv2d(x, w, "VALID")
    gy = DeviceArray.from_host(rng.random(y.shape, dtype=np.float32)); gy.ptr
    dw = y.local_gradients[1][1](gy)
torch.cuda.synchronize()
print("ok", np.asarray(dw).sum())
from oracle import numpy_path as orc
xs = np.asarray(x.array, np.float64); ws = np.asarray(w.array, np.float64); g64 = np.asarray(gy, np.float64)
y0, dx0, dw0 = orc.conv2d_grads(xs[:8], ws, g64[:8], "VALID", (1, 1))
x8 = sp.Variable(xs[:8]); y8 = sp.conv2d(x8, w, "VALID")
print("fprop err", np.abs(np.asarray(y8.array) - y0).max() / np.abs(y0).max())
g8 = DeviceArray.from_host(g64[:8]); g8.ptr
print("wgrad err", np.abs(np.asarray(y8.local_gradients[1][1](g8)) - dw0).max() / np.abs(dw0).max())
for name, fn in (("fprop", lambda: sp.conv2d(x, w, "VALID")), ("wgrad", lambda: y.local_gradients[1][1](gy))):
    for pad in ("VALID", "SAME"):
        if name == "fprop":
            f = (lambda pad=pad: sp.conv2d(x, w, pad))
        else:
            yy = sp.conv2d(x, w, pad); gg = DeviceArray.from_host(rng.random(yy.shape, dtype=np.float32)); gg.ptr
            f = (lambda yy=yy, gg=gg: yy.local_gradients[1][1](gg))
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): f()
        e1.record(); torch.cuda.synchronize()
        print(name, pad, f"{e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call (back-to-back, warm L2)")
