"""conv2d data gradient + maxpool2d/leaky_relu backward: the fused epilogue vs the two launches.
Device time back to back (warm L2).  GPU box only."""
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
rng = np.random.default_rng(0)
U8 = np.dtype(np.uint8)
for name, n, up, cin, cout in (("cifar_dgrad3", 128, 14, 128, 128), ("b400_c256", 400, 14, 256, 64)):
    ph = (up + 1) // 2
    xs = (n, ph, ph, cin)
    w = rng.random((3, 3, cin, cout), dtype=np.float32) - 0.5
    geom = _conv_plan(xs, w.shape, "VALID", (1, 1))[0]
    gy = DeviceArray.from_host(rng.random((n, geom.OH, geom.OW, cout), dtype=np.float32) - 0.5)
    wd = DeviceArray.from_host(w)
    pn = n * ph * ph * cin
    fn_ = n * up * up * cin
    idx = DeviceArray.from_host(rng.integers(0, 4, pn).astype(np.uint8), dtype=np.uint8)
    ypool = DeviceArray.from_host(rng.random(pn, dtype=np.float32) - 0.5)
    xpre = DeviceArray.from_host(rng.random(fn_, dtype=np.float32) - 0.5)
    dpool, dfull, dfull2 = DeviceArray.empty((pn,)), DeviceArray.empty((fn_,)), DeviceArray.empty((fn_,))
    for a in (gy, wd, idx, ypool, xpre): a.ptr
    pg = (n, up, up, cin, 2, 2, 2, 2, 0, 0, ph, ph)
    flag = __import__("ctypes").c_int(0)
    F.sp_conv2d_dgrad_unpool_supported(geom, up, up, 1, flag)
    calls = {
        "dgrad": lambda: F.sp_conv2d_dgrad(gy.ptr, wd.ptr, dpool.ptr, geom, 1, st),
        "pool_bwd": lambda: F.sp_maxpool2d_leaky_bwd(dpool.ptr, idx.ptr, xpre.ptr, dfull.ptr, *pg, 0.02, st),
    }
    if flag.value:
        calls["fused"] = lambda: F.sp_conv2d_dgrad_unpool_leaky(gy.ptr, wd.ptr, idx.ptr, ypool.ptr, dfull2.ptr, geom, up, up, 1, 0.02, st)
    print(name, "supported" if flag.value else "unsupported")
    for cname, fn in calls.items():
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30): fn()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 30)
        print(f"   {cname:10s} b2b {best*1e3:6.1f} us")
