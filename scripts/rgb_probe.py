"""First-layer (C = 3, 3x3) kernels: parity against the float64 oracle and device time of
fprop / wgrad in both precision modes (TF32 = tensor-core filter gradient, FP32 = CUDA-core
kernels) at the CIFAR and VGG first-layer shapes.  Raw C-ABI calls, preallocated buffers;
"b2b" = 30 back-to-back launches in one event pair (warm L2), "cold" = L2 flushed.  GPU box only.
"""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import smallpebble_b200 as sp  # noqa: E402
from bench import load_peaks  # noqa: E402
from oracle import numpy_path as orc  # noqa: E402
from smallpebble_b200 import _abi  # noqa: E402
from smallpebble_b200.array import DeviceArray, stream  # noqa: E402
from smallpebble_b200.core import _conv_plan  # noqa: E402

SHAPES = [
    ("cifar_l1", (128, 32, 32, 3), 32, "VALID"),
    ("vgg_l1", (256, 64, 64, 3), 64, "SAME"),
    ("small_same_k16", (8, 17, 23, 3), 16, "SAME"),
]


def main():
    hbm = load_peaks()["hbm_gbs"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    F = _abi.f
    st = stream()
    rng = np.random.default_rng(0)
    rows = []
    only = os.environ.get("SP_PROBE_ONLY")
    for name, xs, k, pad in SHAPES:
        if only and name not in only.split(","):
            continue
        x = rng.random(xs, dtype=np.float32) - 0.5
        w = rng.random((3, 3, 3, k), dtype=np.float32) - 0.5
        geom = _conv_plan(xs, w.shape, pad, (1, 1))[0]
        oh, ow = geom.OH, geom.OW
        gy = rng.random((xs[0], oh, ow, k), dtype=np.float32) - 0.5
        xd, wd, gd = DeviceArray.from_host(x), DeviceArray.from_host(w), DeviceArray.from_host(gy)
        yd, dwd = DeviceArray.empty(gy.shape), DeviceArray.empty(w.shape)
        y0 = dw0 = None
        if x.size < 2e6:
            y0, _, dw0 = orc.conv2d_grads(x.astype(np.float64), w.astype(np.float64),
                                          gy.astype(np.float64), pad, (1, 1))
        ref_dw = None
        for prec, label in ((0, "fp32"), (1, "tf32")):
            wsb = _abi.load().sp_conv2d_wgrad_workspace(geom, prec)
            wsa = DeviceArray.empty((max(wsb, 4) // 4,))
            calls = {
                "fprop": lambda: F.sp_conv2d_fprop(xd.ptr, wd.ptr, yd.ptr, geom, prec, st),
                "wgrad": lambda: F.sp_conv2d_wgrad(xd.ptr, gd.ptr, dwd.ptr, geom, prec, wsa.ptr, wsb, st),
            }
            for op, fn in calls.items():
                fn()
                torch.cuda.synchronize()
                best = 1e9
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(30):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / 30)
                cold = []
                for _ in range(7):
                    F.sp_l2_flush(flush.data_ptr(), 256 << 20, st)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    torch.cuda.synchronize()
                    cold.append(e0.elapsed_time(e1))
                cold_ms = float(np.median(cold))
                nbytes = 4.0 * (x.size + gy.size + w.size)
                row = {"shape": name, "op": op, "precision": label, "b2b_us": best * 1e3,
                       "cold_us": cold_ms * 1e3, "cold_gbs": nbytes / cold_ms / 1e6,
                       "cold_frac_hbm": nbytes / cold_ms / 1e6 / hbm}
                if op == "wgrad":
                    got = np.asarray(dwd.numpy(), np.float64)
                    if dw0 is not None:
                        row["worst_vs_f64"] = float(np.max(np.abs(got - dw0)) / np.max(np.abs(dw0)))
                    if ref_dw is None:
                        ref_dw = got
                    else:
                        row["worst_vs_fp32_kernel"] = float(np.max(np.abs(got - ref_dw)) / np.max(np.abs(ref_dw)))
                elif y0 is not None:
                    got = np.asarray(yd.numpy(), np.float64)
                    row["worst_vs_f64"] = float(np.max(np.abs(got - y0)) / np.max(np.abs(y0)))
                rows.append(row)
                print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "rgb_probe.json"), "w") as fh:
        json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
