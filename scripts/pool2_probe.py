"""conv2d (RGB 3x3) + leaky_relu + maxpool2d(2,2,s2): the fused kernel vs the two launches.
Device time, back to back (warm L2) and cold (L2 flushed).  GPU box only."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallpebble_b200 as sp
from smallpebble_b200 import _abi
from smallpebble_b200.array import DeviceArray, stream
from smallpebble_b200.core import _conv_plan

F = _abi.f
st = stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rng = np.random.default_rng(0)
xs, k, pad = (128, 32, 32, 3), 32, "SAME"
x = rng.random(xs, dtype=np.float32) - 0.5
w = rng.random((3, 3, 3, k), dtype=np.float32) - 0.5
geom = _conv_plan(xs, w.shape, pad, (1, 1))[0]
oh, ow = geom.OH, geom.OW
xd, wd = DeviceArray.from_host(x), DeviceArray.from_host(w)
yd = DeviceArray.empty((xs[0], oh, ow, k))
pn = xs[0] * (oh // 2) * (ow // 2) * k
yp, yp2 = DeviceArray.empty((pn,)), DeviceArray.empty((pn,))
idx, idx2 = DeviceArray.empty((pn,), np.dtype(np.uint8), pn), DeviceArray.empty((pn,), np.dtype(np.uint8), pn)
spl, spl2 = DeviceArray.empty((2 * pn,)), DeviceArray.empty((2 * pn,))
pg = (xs[0], oh, ow, k, 2, 2, 2, 2, 0, 0, oh // 2, ow // 2)
X3 = _abi.PRECISION_TF32X3
calls = {
    "conv": lambda: F.sp_conv2d_fprop(xd.ptr, wd.ptr, yd.ptr, geom, X3, st),
    "pool_split": lambda: F.sp_maxpool2d_fwd_split(yd.ptr, yp.ptr, idx.ptr, spl.ptr, 32, 0, *pg, 1, 0.02, st),
    "fused": lambda: F.sp_conv2d_leaky_pool2_fwd(xd.ptr, wd.ptr, yp2.ptr, idx2.ptr, spl2.ptr, geom, X3, 0.02, st),
    "fused_nosplit": lambda: F.sp_conv2d_leaky_pool2_fwd(xd.ptr, wd.ptr, yp2.ptr, idx2.ptr, None, geom, X3, 0.02, st),
}
for name, fn in calls.items():
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 30)
    cold = []
    for _ in range(7):
        F.sp_l2_flush(flush.data_ptr(), 256 << 20, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        cold.append(e0.elapsed_time(e1))
    print(f"{name:14s} b2b {best*1e3:6.1f} us   cold {sorted(cold)[3]*1e3:6.1f} us")
calls["fused"]()
torch.cuda.synchronize()
print("equal:", np.array_equal(np.asarray(yp), np.asarray(yp2)), np.array_equal(np.asarray(idx), np.asarray(idx2)),
      np.array_equal(np.asarray(spl), np.asarray(spl2)))
