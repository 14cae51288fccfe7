"""Host->device copy time of one float64 CIFAR batch (3.1 MB, pinned), one stream vs split over two."""
import time, torch, numpy as np
n = 128 * 32 * 32 * 3
src = torch.empty(n, dtype=torch.float64).pin_memory()
src.uniform_()
dst = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, reps=50):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2] * 1e6
def one():
    with torch.cuda.stream(s1): dst.copy_(src, non_blocking=True)
def two():
    h = n // 2
    with torch.cuda.stream(s1): dst[:h].copy_(src[:h], non_blocking=True)
    with torch.cuda.stream(s2): dst[h:].copy_(src[h:], non_blocking=True)
def ev(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        e0.record(); 
        for _ in range(20): dst.copy_(src, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3
print(f"one stream {timeit(one):.1f} us wall, two streams {timeit(two):.1f} us wall, event-timed b2b {ev(one):.1f} us "
      f"({n * 8 / 1e3 / ev(one):.1f} GB/s)")
f32 = torch.empty(n, dtype=torch.float32).pin_memory(); d32 = torch.empty(n, dtype=torch.float32, device="cuda")
def half():
    with torch.cuda.stream(s1): d32.copy_(f32, non_blocking=True)
print(f"fp32 batch (1.5 MB) {timeit(half):.1f} us wall")
a = src.numpy(); b = f32.numpy()
t0 = time.perf_counter()
for _ in range(20): np.copyto(b, a, casting="same_kind")
print(f"numpy float64->float32 narrowing on one host thread: {(time.perf_counter() - t0) / 20 * 1e6:.0f} us")
