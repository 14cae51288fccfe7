"""README MLP layers: device time of the one-launch linear layer / classifier head kernels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smallpebble_b200 as sp
from smallpebble_b200 import _abi
from smallpebble_b200.array import DeviceArray, stream
F = _abi.f
st = stream()
rng = np.random.default_rng(0)
def dev(*shape): return DeviceArray.from_host(rng.random(shape, dtype=np.float32) - 0.5)
def timeit(fn):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 50)
    return best * 1e3
for m, k, n in ((200, 784, 100), (200, 100, 100), (128, 1152, 100)):
    a, b, bias, c = dev(m, k), dev(k, n), dev(n), DeviceArray.empty((m, n))
    for x in (a, b, bias): x.ptr
    t1 = timeit(lambda: F.sp_gemm_bias_leaky(m, n, k, a.ptr, b.ptr, bias.ptr, c.ptr, 0, 0.02, st))
    t2 = timeit(lambda: F.sp_gemm(0, 0, m, n, k, a.ptr, k, 0, b.ptr, n, 0, c.ptr, n, m * n, 1, 0, st))
    t3 = timeit(lambda: F.sp_gemm(0, 0, m, n, k, a.ptr, k, 0, b.ptr, n, 0, c.ptr, n, m * n, 1, 1, st))
    print(f"[{m}x{k}]x[{k}x{n}]: gemm_bias_leaky {t1:.1f} us   sp_gemm fp32 {t2:.1f} us   sp_gemm tf32 {t3:.1f} us")
m, k, n = 200, 100, 10
a, b, bias = dev(m, k), dev(k, n), dev(n)
lab = DeviceArray.from_host(rng.integers(0, n, m), dtype=np.int32)
lg, pr, dl = DeviceArray.empty((m, n)), DeviceArray.empty((m, n)), DeviceArray.empty((m, n))
loss, cs = DeviceArray.empty(()), DeviceArray.empty((n,))
for x in (a, b, bias, lab): x.ptr
t = timeit(lambda: F.sp_gemm_bias_softmax_xent(m, n, k, a.ptr, b.ptr, bias.ptr, lab.ptr, lg.ptr, pr.ptr, loss.ptr, dl.ptr, cs.ptr, 0, st))
print(f"head [{m}x{k}]x[{k}x{n}] + softmax-xent: {t:.1f} us")
z = DeviceArray.empty((4,))
t = timeit(lambda: F.sp_fill_f32(z.ptr, 4, 1.0, st))
print(f"empty-ish kernel (fill 4 floats): {t:.1f} us")
