"""In-kernel timeline of the 2-SM TMA conv kernel (cluster 0, leader CTA), from clock64
stamps written through sp_debug_tma2_trace.  Slots (int64):
  0 kernel entry | 1 setup done (barriers, TMEM, cluster sync) | 2..5 MMA issuer: all MMAs of
  tile i issued | 6..9 epilogue: accumulator i ready | 10..13 epilogue: tile i stored |
  14 teardown done | 16+k A producer issues k-block k | 16+96+k B producer | 16+192+k MMA
  issuer sees k-block k landed | 450+6b .. 455+6b (b = 0, 1): pooled epilogue (EPI = 3) of the
  first tile, column block b: start, TMEM loaded, parked, barrier 1, pooled, barrier 2
  (scripts/pool_trace.py).

    python scripts/tma2_trace.py [case ...]
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

POLICY = 32 + (128 if os.environ.get("SP_PROBE_HALO") == "1" else 256)
CASES = {
    "vgg_conv1b": ((256, 64, 64, 64), (3, 3, 64, 64), "SAME"),
    "vgg_conv2b": ((256, 32, 32, 128), (3, 3, 128, 128), "SAME"),
    "cnn_conv2": ((128, 15, 15, 32), (3, 3, 32, 128), "VALID"),
    "cnn_conv3": ((128, 7, 7, 128), (3, 3, 128, 128), "VALID"),
    "vgg_conv3b": ((256, 16, 16, 256), (3, 3, 256, 256), "SAME"),
}
KB = 96


def main():
    import torch

    import smallpebble_b200 as sp
    from smallpebble_b200 import _abi, core
    from smallpebble_b200.array import DeviceArray

    sp.set_precision("tf32")
    core.debug_set_kernel_policy(POLICY)
    rng = np.random.default_rng(0)
    buf = torch.zeros(512, dtype=torch.int64, device="cuda")
    names = sys.argv[1:] or list(CASES)
    for name in names:
        xs, ws, pad = CASES[name]
        x = sp.Variable(rng.random(xs, dtype=np.float32) - 0.5)
        w = sp.Variable(rng.random(ws, dtype=np.float32) - 0.5)
        y = sp.conv2d(x, w, pad, (1, 1))
        gy = DeviceArray.from_host(rng.random(y.shape, dtype=np.float32) - 0.5)
        gy.ptr
        (_, fdx), (_, fdw) = y.local_gradients
        for op, fn in (("fprop", lambda: sp.conv2d(x, w, pad, (1, 1))), ("dgrad", lambda: fdx(gy)),
                       ("wgrad", lambda: fdw(gy))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            buf.zero_()
            _abi.f.sp_debug_tma2_trace(buf.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            _abi.f.sp_debug_tma2_trace(None)
            t = buf.cpu().numpy().astype(np.int64)
            t0 = t[0]
            rel = lambda v: int(v - t0) if v else -1  # noqa: E731
            a = [rel(v) for v in t[16:16 + KB] if v]
            b = [rel(v) for v in t[16 + KB:16 + 2 * KB] if v]
            m = [rel(v) for v in t[16 + 2 * KB:16 + 3 * KB] if v]
            print(f"== {name}/{op}: event {1e3 * e0.elapsed_time(e1):.1f} us (with launch); "
                  f"cycles from entry: setup {rel(t[1])} | mma issued {[rel(v) for v in t[2:6]]} | "
                  f"acc ready {[rel(v) for v in t[6:10]]} | stored {[rel(v) for v in t[10:14]]} | "
                  f"end {rel(t[14])}")
            print("   A issue :", a[:40])
            print("   B issue :", b[:40])
            print("   landed  :", m[:40])
            fine = t[352:352 + 90].reshape(30, 3)
            if fine[:, 1].any():
                print("   MMA thread per k-block (cycles): wait-done -> 4 MMAs issued -> commit issued:")
                print("     ", [(int(m_ - l_), int(c_ - m_)) for l_, m_, c_ in
                                 zip(t[16 + 2 * KB:16 + 2 * KB + 30], fine[:, 1], fine[:, 2]) if m_][:20])
            strips = [rel(v) for v in t[320:352] if v]
            if strips:
                print("   strip landed (halo):", strips[:16])
            if len(m) > 8:
                d = np.diff(m)
                print(f"   landed cadence: median {int(np.median(d))} cyc, first {m[0]}, "
                      f"A->landed first {m[0] - a[0]}")


if __name__ == "__main__":
    main()
