"""Host-side cost of one README training step WITHOUT a GPU: the C-ABI functions are replaced
by no-ops and device buffers by CPU byte tensors, so what remains is exactly the Python work the
library does per step (tape construction, get_gradients, plan caches, buffer recycling).  A
profiling harness for this container only -- nothing here is a product path, no numerics run.

    python scripts/host_profile_nogpu.py [steps]
"""

import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from smallpebble_b200 import _abi, array as A  # noqa: E402


def _noop_getattr(self, name):
    if name.startswith("_") or name not in _abi.SIGNATURES:
        raise AttributeError(name)
    fn = (lambda *a: 0)
    setattr(self, name, fn)
    return fn


class _FakeLib:
    def __getattr__(self, name):
        if name == "sp_conv2d_wgrad_workspace":
            return lambda geom, prec: 1 << 20
        return lambda *a: 0


_abi._Functions.__getattr__ = _noop_getattr
_abi.load = lambda: _FakeLib()
A._torch = torch
A._device = torch.device("cpu")
A._raw_stream = lambda idx: 0
A.require_cuda = lambda: A._device
A._keep_alive_until_copied = lambda host: None

import smallpebble_b200 as sp  # noqa: E402
from smallpebble_b200 import core, workloads  # noqa: E402

x, y = workloads.synthetic_batch("cnn", 128)
wl = workloads.cifar_cnn(seed=0)
adam = sp.Adam()
x_dev = A.DeviceArray.empty(x.shape)
y_dev = A.DeviceArray.empty(y.shape, np.int32)


def step():
    wl.x_in.assign_value(sp.Variable(x_dev))
    wl.y_true.assign_value(y_dev)
    loss_val = wl.loss.run()
    gradients = sp.get_gradients(loss_val)
    adam.training_step(wl.learnables, gradients)
    return loss_val


n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for _ in range(50):
    step()
runs = []
for _ in range(12):
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    runs.append((time.perf_counter() - t0) / n)
print(f"python-only host time: min {min(runs) * 1e6:.1f} us/step, median {np.median(runs) * 1e6:.1f}")

if os.environ.get("SP_PROFILE", "1") == "1":
    prof = cProfile.Profile()
    prof.enable()
    for _ in range(300):
        step()
    prof.disable()
    st = pstats.Stats(prof)
    st.sort_stats("tottime").print_stats(35)
