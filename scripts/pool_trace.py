"""In-kernel timeline (cluster 0, leader CTA; sp_debug_tma2_trace) of the README CNN's second
conv layer: bare conv (pool as a separate launch) vs conv + leaky_relu + pool in the epilogue."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallpebble_b200 as sp
from smallpebble_b200 import _abi, core, steady
steady.set_enabled(False)
rng = np.random.default_rng(0)
x = sp.Variable(rng.random((128, 32, 32, 3)) - 0.5)
w1 = sp.Variable((rng.random((3, 3, 3, 32)) - 0.5) * 0.4)
w2 = sp.Variable((rng.random((3, 3, 32, 128)) - 0.5) * 0.1)
w3 = sp.Variable((rng.random((3, 3, 128, 128)) - 0.5) * 0.05)
def fwd():
    h = sp.maxpool2d(sp.leaky_relu(sp.conv2d(x, w1, "VALID")), 2, 2, strides=[2, 2])
    h = sp.maxpool2d(sp.leaky_relu(sp.conv2d(h, w2, "VALID")), 2, 2, strides=[2, 2])
    h = sp.conv2d(h, w3, "VALID")  # consumer: makes the pool above emit the (r, l) split
    h.array.ptr
    return h
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
for fuse in (False, True):
    core._fuse_conv_pool = fuse
    for _ in range(4): fwd()
    torch.cuda.synchronize()
    # trace only the second conv: stop tracing before the third launches
    orig = _abi.f.sp_conv2d_fprop
    state = {"n": 0}
    def spy(*a):
        state["n"] += 1
        if state["n"] == 2 and not fuse: _abi.f.sp_debug_tma2_trace(buf.data_ptr())
        r = orig(*a)
        if state["n"] == 2 and not fuse: _abi.f.sp_debug_tma2_trace(None)
        return r
    origp = _abi.f.sp_conv2d_leaky_pool2_fwd
    statep = {"n": 0}
    def spyp(*a):
        statep["n"] += 1
        if statep["n"] == 2: _abi.f.sp_debug_tma2_trace(buf.data_ptr())
        r = origp(*a)
        if statep["n"] == 2: _abi.f.sp_debug_tma2_trace(None)
        return r
    buf.zero_()
    _abi.f.sp_conv2d_fprop = spy; _abi.f.sp_conv2d_leaky_pool2_fwd = spyp
    try:
        fwd(); torch.cuda.synchronize()
    finally:
        _abi.f.sp_conv2d_fprop = orig; _abi.f.sp_conv2d_leaky_pool2_fwd = origp
    t = buf.cpu().numpy().astype(np.int64)
    t0 = t[0]
    rel = lambda v: int(v - t0) if v else -1
    if fuse:
        print("   first tile, column blocks 0 / 1: [start, tmem loaded, parked, barrier 1, pooled, barrier 2]:",
              [rel(v) for v in t[450:456]], [rel(v) for v in t[456:462]])
    print(f"fused={fuse}: setup {rel(t[1])} | mma issued {[rel(v) for v in t[2:6]]} | acc ready {[rel(v) for v in t[6:10]]} | stored {[rel(v) for v in t[10:14]]} | end {rel(t[14])}")
core._fuse_conv_pool = True
