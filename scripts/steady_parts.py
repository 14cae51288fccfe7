"""Device time of the three steady-state graphs (forward / backward / update) replayed alone,
and of the eager kernels of each phase back to back."""
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
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    sp.distributed.init()
x, y = workloads.synthetic_batch(kind, batch)
wl = workloads.BUILDERS[kind](seed=0)
adam = sp.Adam()
xd = DeviceArray.from_host(x); yd = DeviceArray.from_host(y, dtype=np.int32); xd.ptr; yd.ptr
for i in range(12):
    wl.x_in.assign_value(sp.Variable(xd)); wl.y_true.assign_value(yd)
    lv = wl.loss.run(); g = sp.get_gradients(lv); adam.training_step(wl.learnables, g)
torch.cuda.synchronize()
rec = g._steady[0] if g._steady else None
assert rec is not None, steady.stats
def timeit(graph, n=200):
    for _ in range(5): graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): graph.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
tf_, tb_, tu_ = timeit(rec.graph_f), timeit(rec.graph_b), timeit(rec.graph_u)  # every rank
if int(os.environ.get("RANK", "0")) == 0:
    print(f"{kind} b{batch} {sp.get_precision()}: F {tf_:.1f} us ({rec.kernels_f} calls)  "
          f"B {tb_:.1f} us ({rec.kernels_b} calls)  U {tu_:.1f} us ({rec.kernels_u} calls)")

import time
def loop(n=300):
    for i in range(20):
        wl.x_in.assign_value(sp.Variable(xd)); wl.y_true.assign_value(yd)
        lv = wl.loss.run(); adam.training_step(wl.learnables, sp.get_gradients(lv))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    tf = tb = tu = 0.0
    for i in range(n):
        wl.x_in.assign_value(sp.Variable(xd)); wl.y_true.assign_value(yd)
        a = time.perf_counter(); lv = wl.loss.run(); b = time.perf_counter()
        g = sp.get_gradients(lv); c = time.perf_counter()
        adam.training_step(wl.learnables, g); d = time.perf_counter()
        tf += b - a; tb += c - b; tu += d - c
    host = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, host / n * 1e6, tf / n * 1e6, tb / n * 1e6, tu / n * 1e6
r = loop()
if int(os.environ.get("RANK", "0")) == 0:
    print("loop: %.1f us/step device, host %.1f us/step (run %.1f, get_gradients %.1f, adam %.1f)" % r)
if os.environ.get("SP_PEER_TRACE") == "1" and sp.distributed._peer["handle"]:
    import ctypes as C
    from smallpebble_b200 import _abi
    torch.cuda.synchronize()
    out = (C.c_uint64 * 8)()
    _abi.load().sp_debug_peer_trace(sp.distributed._peer["handle"], out)
    t = list(out)
    print("rank", os.environ.get("RANK"), "peer kernel block 0 (us from entry): bucket written %.1f | fenced %.1f | "
          "flags seen %.1f | update done %.1f" % tuple((t[i] - t[0]) / 1e3 for i in (1, 2, 3, 4)), flush=True)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    sp.distributed.barrier()
    os._exit(0)
