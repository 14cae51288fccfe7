"""In-stream device time of every kernel-launching ABI call of one training step.

    python scripts/step_profile.py [cnn|mlp|vgg] [batch] [precision]

Every launching call of `reps` steps is bracketed by a CUDA event pair while the device is
kept busy ahead of the host (so the pairs contain device time, not host gaps).  Event pairs
break the programmatic-dependent-launch overlap and add ~1-2 us each: read the numbers as
shares of the step; the un-instrumented eager and graph-replayed step times are printed
beside them.  (ncu launch lists under profiles/ are cold-cache and serialised.)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import smallpebble_b200 as sp  # noqa: E402
from smallpebble_b200 import _abi, workloads  # noqa: E402
from smallpebble_b200.array import DeviceArray  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cnn"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else {"cnn": 128, "mlp": 200, "vgg": 256}[kind]
if len(sys.argv) > 3:
    sp.set_precision(sys.argv[3])
reps = 6

x, y = workloads.synthetic_batch(kind, batch)
wl = workloads.BUILDERS[kind](seed=0)
adam = sp.Adam()
xs = [DeviceArray.from_host(np.random.default_rng(i).random(x.shape)) for i in range(4)]
yd = DeviceArray.from_host(y, dtype=np.int32)
[a.ptr for a in xs], yd.ptr
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def step(i=0):
    wl.x_in.assign_value(sp.Variable(xs[i % 4]))
    wl.y_true.assign_value(yd)
    lv = wl.loss.run()
    adam.training_step(wl.learnables, sp.get_gradients(lv))


def describe(name, a):
    if name.startswith("sp_conv2d"):
        g = next(v for v in a if hasattr(v, "contents") or hasattr(v, "KH"))
        g = g.contents if hasattr(g, "contents") else g
        return f"N{g.N} {g.H}x{g.W} C{g.C}->K{g.K} k{g.KH} s{g.sy} prec={[v for v in a if isinstance(v, int)][:1]}"
    if name in ("sp_gemm", "sp_gemm_bias"):
        ints = [v for v in a if isinstance(v, int)][:5]
        return f"{ints}"
    return ""


for i in range(30):
    step(i)
torch.cuda.synchronize()

# plain eager / graph step times
for label, fn in (("eager", step),):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200 if kind != "vgg" else 20
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        fn(i)
    host = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/step device-timed, host issue "
          f"{host / n * 1e6:.1f} us/step")

_abi.prof.records.clear()
_abi.prof.select = None
_abi.prof.active = True
step()
torch.cuda.synchronize()
per_seq = {}
for r in range(reps):
    _abi.prof.records.clear()
    for _ in range(6):
        big.zero_()
    _abi.prof.begin_step()
    step(r)
    torch.cuda.synchronize()
    recs = list(_abi.prof.records)
    for name, seq, a, e0, e1 in recs:
        per_seq.setdefault(seq, [name, describe(name, a), []])[2].append(e0.elapsed_time(e1) * 1e3)
    if r == reps - 1:
        span = recs[0][3].elapsed_time(recs[-1][4]) * 1e3
_abi.prof.active = False
total = 0.0
print(f"{'seq':>3} {'call':34s} {'us(min)':>8} {'us(med)':>8}  detail")
for seq in sorted(per_seq):
    name, det, ts = per_seq[seq]
    total += min(ts)
    print(f"{seq:3d} {name:34s} {min(ts):8.1f} {float(np.median(ts)):8.1f}  {det}")
print(f"sum of minima {total:.1f} us over {len(per_seq)} calls; instrumented span of the last step {span:.1f} us")

try:
    gstep = wl.captured_train_step(adam)
    for i in range(3):
        gstep(xs[i % 4], yd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200 if kind != "vgg" else 20
    e0.record()
    for i in range(n):
        gstep(xs[i % 4], yd)
    e1.record()
    torch.cuda.synchronize()
    print(f"graph replay: {e0.elapsed_time(e1) / n * 1e3:.1f} us/step")
except Exception as exc:  # noqa: BLE001
    print("graph replay failed:", exc)
