"""Where does the host time of one training step go?  Runs the CIFAR CNN step many times
without device syncs, reports wall time per step (CPU-bound rate), a cProfile of the Python
side, and micro-timings of the upload paths.  GPU box only."""

import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import smallpebble_b200 as sp  # noqa: E402
from smallpebble_b200 import workloads  # noqa: E402
from smallpebble_b200.array import DeviceArray  # noqa: E402

x, y = workloads.synthetic_batch("cnn", 128)
wl = workloads.cifar_cnn(seed=0)
adam = sp.Adam()
x_dev = DeviceArray.from_host(x)
y_dev = DeviceArray.from_host(y, dtype=np.int32)
x_dev.ptr, y_dev.ptr


def step():
    wl.x_in.assign_value(sp.Variable(x_dev))
    wl.y_true.assign_value(y_dev)
    loss_val = wl.loss.run()
    gradients = sp.get_gradients(loss_val)
    adam.training_step(wl.learnables, gradients)
    return loss_val


for _ in range(10):
    step()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    step()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_total = time.perf_counter() - t0
print(f"resident step: host issue {1e6 * t_issue / n:.1f} us/step, with final sync "
      f"{1e6 * t_total / n:.1f} us/step")

prof = cProfile.Profile()
prof.enable()
for _ in range(100):
    step()
prof.disable()
torch.cuda.synchronize()
st = pstats.Stats(prof)
st.sort_stats("tottime").print_stats(22)

# upload micro-benchmarks
for name, arr in (("float64 batch", x), ("float32 batch", x.astype(np.float32))):
    for _ in range(3):
        DeviceArray.from_host(arr).ptr
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        d = DeviceArray.from_host(arr)
    t1 = time.perf_counter()
    for _ in range(50):
        d = DeviceArray.from_host(arr)
        d.ptr
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: from_host {1e6 * (t1 - t0) / 50:.0f} us, from_host+upload {1e6 * (t2 - t1) / 50:.0f} us")
t0 = time.perf_counter()
for _ in range(50):
    a32 = np.array(x, dtype=np.float32, order="C", copy=True)
t1 = time.perf_counter()
print(f"np f64->f32 convert: {1e6 * (t1 - t0) / 50:.0f} us")
loss = step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    float(loss.array)
print(f"scalar D2H read: {1e6 * (time.perf_counter() - t0) / 50:.1f} us")

# ---- where the HOST time of a step goes: C-ABI calls (driver launch included) vs Python --------
# Each step is followed by a sync so the launch queue never fills: the per-step wall time up to
# the last call is pure host issue time.
from smallpebble_b200 import _abi  # noqa: E402

acc = {}


def _timed(name, fn):
    def call(*a):
        t = time.perf_counter_ns()
        fn(*a)
        d = time.perf_counter_ns() - t
        e = acc.setdefault(name, [0, 0])
        e[0] += d
        e[1] += 1
    return call


torch.cuda.synchronize()
n = 200
host_ns = 0
for _ in range(n):
    t0 = time.perf_counter_ns()
    step()
    host_ns += time.perf_counter_ns() - t0
    torch.cuda.synchronize()
print(f"host issue time, device idle at step start: {host_ns / n / 1e3:.1f} us/step")
for name in [k for k in _abi.f.__dict__ if k in _abi.SIGNATURES]:
    setattr(_abi.f, name, _timed(name, getattr(_abi.f, name)))
host_ns = 0
for _ in range(n):
    t0 = time.perf_counter_ns()
    step()
    host_ns += time.perf_counter_ns() - t0
    torch.cuda.synchronize()
total_abi = sum(v[0] for v in acc.values())
print(f"with call timers: {host_ns / n / 1e3:.1f} us/step; inside C-ABI calls {total_abi / n / 1e3:.1f} us/step "
      f"({sum(v[1] for v in acc.values()) / n:.0f} calls/step)")
for name, (ns, cnt) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"   {name:28s} {cnt / n:5.1f} calls/step  {ns / cnt / 1e3:6.2f} us/call  {ns / n / 1e3:6.1f} us/step")
