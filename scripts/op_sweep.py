"""Op-level sweep (BASELINE.json configs[4]): conv2d fprop / dgrad / wgrad, maxpool2d,
leaky_relu, bias add, Adam at the CIFAR-CNN and VGG-6 layer shapes, timed with CUDA events on
the launch stream (L2 flushed between repeats), reported against the measured B200 rooflines.

    python scripts/op_sweep.py [--precision tf32|fp32] [--reps 20] [--out gpurun_out/op_sweep.json]
"""

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "op_sweep.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()

    import torch

    import smallpebble_b200 as sp
    from bench import load_peaks
    from smallpebble_b200 import _abi
    from smallpebble_b200.array import DeviceArray, stream

    sp.set_precision(args.precision)
    peaks = load_peaks()
    tensor_peak = peaks["bf16_tflops"] / 2 if args.precision == "tf32" else 74.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rng = np.random.default_rng(0)

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            _abi.f.sp_l2_flush(flush.data_ptr(), 256 << 20, stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        return float(np.median(times)), float(np.min(times))

    rows = []

    def record(name, kind, work, ms_med, ms_min):
        if kind == "tensor":
            ach = work / (ms_med * 1e-3) / 1e12
            row = {"op": name, "bound": "tensor", "ms": ms_med, "ms_min": ms_min,
                   "achieved_tflops": ach, "peak_tflops": tensor_peak, "frac": ach / tensor_peak}
        else:
            ach = work / (ms_med * 1e-3) / 1e9
            row = {"op": name, "bound": "hbm", "ms": ms_med, "ms_min": ms_min,
                   "achieved_gbs": ach, "peak_gbs": peaks["hbm_gbs"], "frac": ach / peaks["hbm_gbs"]}
        rows.append(row)
        print(json.dumps(row), flush=True)

    conv_cfgs = [
        ("cnn_conv1", (128, 32, 32, 3), (3, 3, 3, 32), "VALID"),
        ("cnn_conv2", (128, 15, 15, 32), (3, 3, 32, 128), "VALID"),
        ("cnn_conv3", (128, 7, 7, 128), (3, 3, 128, 128), "VALID"),
        ("vgg_conv1b_64_64", (256, 64, 64, 64), (3, 3, 64, 64), "SAME"),
        ("vgg_conv2a_64_128", (256, 32, 32, 64), (3, 3, 64, 128), "SAME"),
        ("vgg_conv2b_128_128", (256, 32, 32, 128), (3, 3, 128, 128), "SAME"),
        ("vgg_conv3a_128_256", (256, 16, 16, 128), (3, 3, 128, 256), "SAME"),
        ("vgg_conv3b_256_256", (256, 16, 16, 256), (3, 3, 256, 256), "SAME"),
        ("sweep_n64_c512_k1", (64, 32, 32, 512), (1, 1, 512, 512), "SAME"),
        ("sweep_n64_c64_k5_s2", (64, 32, 32, 64), (5, 5, 64, 64), "SAME"),
    ]
    if args.quick:
        conv_cfgs = conv_cfgs[:4]
    for name, xs, ws, pad in conv_cfgs:
        strides = (2, 2) if name.endswith("_s2") else (1, 1)
        x = sp.Variable((rng.random(xs, dtype=np.float32) - 0.5))
        w = sp.Variable((rng.random(ws, dtype=np.float32) - 0.5))
        y = sp.conv2d(x, w, pad, strides)
        gy = DeviceArray.from_host(rng.random(y.shape, dtype=np.float32) - 0.5)
        gy.ptr
        (_, fdx), (_, fdw) = y.local_gradients
        flops = 2.0 * np.prod(y.shape) * ws[0] * ws[1] * ws[2]
        record(f"{name}/fprop", "tensor", flops, *timeit(lambda: sp.conv2d(x, w, pad, strides)))
        record(f"{name}/dgrad", "tensor", flops, *timeit(lambda: fdx(gy)))
        record(f"{name}/wgrad", "tensor", flops, *timeit(lambda: fdw(gy)))

    for name, xs in [("cnn_pool1", (128, 30, 30, 32)), ("vgg_pool1", (256, 64, 64, 64)),
                     ("vgg_pool2", (256, 32, 32, 128))]:
        x = sp.Variable(rng.random(xs, dtype=np.float32) - 0.5)
        y = sp.maxpool2d(x, 2, 2, strides=[2, 2])
        gy = DeviceArray.from_host(rng.random(y.shape, dtype=np.float32))
        gy.ptr
        ins, outs = int(np.prod(xs)), int(np.prod(y.shape))
        record(f"{name}/maxpool_fwd", "hbm", 4.0 * ins + 5.0 * outs,
               *timeit(lambda: sp.maxpool2d(x, 2, 2, strides=[2, 2])))
        record(f"{name}/maxpool_bwd", "hbm", 4.0 * ins + 5.0 * outs,
               *timeit(lambda: y.local_gradients[0][1](gy)))
        z = sp.leaky_relu(x)
        gz = DeviceArray.from_host(rng.random(xs, dtype=np.float32))
        gz.ptr
        record(f"{name}/leaky_relu_fwd", "hbm", 8.0 * ins, *timeit(lambda: sp.leaky_relu(x)))
        record(f"{name}/leaky_relu_bwd", "hbm", 12.0 * ins,
               *timeit(lambda: z.local_gradients[0][1](gz)))

    n = 16_000_000
    p = sp.Variable(rng.random(n, dtype=np.float32))
    g = DeviceArray.from_host(rng.random(n, dtype=np.float32))
    g.ptr
    adam = sp.Adam()
    record("adam_16M_params", "hbm", 28.0 * n, *timeit(lambda: adam.training_step([p], {p: g})))
    a = sp.Variable(rng.random((65536, 256), dtype=np.float32))
    b = sp.Variable(rng.random((256,), dtype=np.float32))
    out = sp.add(a, b)
    record("bias_add_65536x256", "hbm", 8.0 * a.array.size, *timeit(lambda: sp.add(a, b)))
    gb = DeviceArray.from_host(rng.random(out.shape, dtype=np.float32))
    gb.ptr
    record("bias_grad_colsum_65536x256", "hbm", 4.0 * a.array.size,
           *timeit(lambda: out.local_gradients[1][1](gb)))

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump({"precision": args.precision, "peaks": peaks, "tensor_peak_tflops": tensor_peak,
                   "tensor_peak_note": "TF32 dense taken as 1/2 of the measured bf16 burst peak",
                   "timing": "median of CUDA-event timings on the launch stream, L2 flushed "
                             "(256 MiB memset) before every repeat", "rows": rows}, fh, indent=1)


if __name__ == "__main__":
    main()
