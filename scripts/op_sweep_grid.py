"""BASELINE.json configs[4] as specified: conv2d fprop / dgrad / wgrad and maxpool2d over the grid

    batch N in {32, 128, 512, 1024}  x  channels C = K in {16, 64, 256, 512}
    x  kernel 1 / 3 / 5  x  stride 1 / 2          (32 x 32 images, SAME padding)

every point (a) checked against an INDEPENDENT float64 implementation -- torch's CPU conv2d /
max_pool autograd, which SURVEY.md section 4 shows equal to the reference to 1e-12 -- on the
full-size launch (the per-image ops are compared on the first images of the batch; the filter
gradient, a sum over the batch, through linearity: an output gradient that is zero outside
the first images); (b) timed with CUDA events on the launch stream, L2 flushed before every
repeat, straight through the C ABI (sp_conv2d_fprop / dgrad / wgrad, sp_maxpool2d_fwd / bwd),
against the rooflines: tensor pipe (measured TF32 peak) and HBM (algorithmic bytes); (c) with
the NumPy restatement of the reference (oracle/numpy_path.py, float64, the host's cores) timed
beside it where its patch matrix fits in 2 GiB.

    python scripts/op_sweep_grid.py [--quick] [--out gpurun_out/r02_op_sweep_grid.json]
"""

import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tf32_peak(peaks):
    """(roofline, note, cuBLAS TF32 figure or None).  The roofline is the LARGER of the two
    measured candidates -- 1/2 x bf16 burst (MEASURED_PEAKS.json) and the calibrated cuBLAS TF32
    GEMM (scripts/tf32_peak.py) -- see DESIGN.md section 3."""
    half = peaks["bf16_tflops"] / 2
    for cand in ("profiles/r02_tf32_peak.json", "gpurun_out/r02_tf32_peak.json"):
        path = os.path.join(ROOT, cand)
        if os.path.exists(path):
            with open(path) as fh:
                cublas = json.load(fh)["tf32_tflops"]
            if cublas > half:
                return cublas, f"measured cuBLAS TF32 GEMM burst ({cand})", cublas
            return half, ("1/2 x measured bf16 burst (MEASURED_PEAKS.json); the calibrated cuBLAS "
                          f"TF32 GEMM reaches {cublas:.0f} ({cand})"), cublas
    return half, "1/2 x measured bf16 burst (MEASURED_PEAKS.json)", None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-numpy", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_op_sweep_grid.json"))
    args = ap.parse_args()

    import torch
    import torch.nn.functional as TF

    import smallpebble_b200 as sp
    from bench import load_peaks
    from oracle import numpy_path as orc  # the CPU leg beside the GPU numbers (checker / baseline)
    from smallpebble_b200 import _abi, core
    from smallpebble_b200.array import FLOAT, DeviceArray, stream

    F = _abi.f
    sp.set_precision("tf32x1")  # single-pass TF32: the tensor-pipe roofline is quoted on this
    peaks = load_peaks()
    tpeak, tpeak_note, cublas_tf32 = tf32_peak(peaks)
    hbm = peaks["hbm_gbs"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(0)

    def dev(shape, scale=1.0):
        t = (torch.rand(shape, device="cuda", generator=gen) - 0.5) * scale
        return t, DeviceArray(tuple(shape), FLOAT, t, t.data_ptr())

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            F.sp_l2_flush(flush.data_ptr(), 256 << 20, stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    def err(got, ref):
        ref = ref.double()
        return float((got.double().cpu() - ref).abs().max() / max(float(ref.abs().max()), 1e-30))

    Ns = [32, 128, 512, 1024]
    Cs = [16, 64, 256, 512]
    if args.quick:
        Ns, Cs = [32, 128], [16, 64]
    H = W = 32
    rows = []
    t_start = time.time()
    for N in Ns:
        for C in Cs:
            K = C
            xt, xd = dev((N, H, W, C))
            for k in (1, 3, 5):
                wt, wd = dev((k, k, C, K), 2.0 / math.sqrt(k * k * C))
                for s in (1, 2):
                    plan = core._conv_plan((N, H, W, C), (k, k, C, K), "SAME", (s, s))
                    geom, (_, OH, OW, _) = plan[0], plan[1]
                    yt, yd = dev((N, OH, OW, K))
                    yt.zero_()
                    gt, gd = dev((N, OH, OW, K))
                    sub = min(N, 2)
                    gt[sub:] = 0  # filter gradient = that of the first `sub` images (linearity)
                    dxt, dxd = dev((N, H, W, C))
                    dwt, dwd = dev((k, k, C, K))
                    ws_bytes = _abi.load().sp_conv2d_wgrad_workspace(geom, _abi.PRECISION_TF32)
                    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device="cuda")
                    prec = _abi.PRECISION_TF32
                    st = stream()
                    ops = {
                        "fprop": lambda: F.sp_conv2d_fprop(xd.ptr, wd.ptr, yd.ptr, geom, prec, st),
                        "dgrad": lambda: F.sp_conv2d_dgrad(gd.ptr, wd.ptr, dxd.ptr, geom, prec, st),
                        "wgrad": lambda: F.sp_conv2d_wgrad(xd.ptr, gd.ptr, dwd.ptr, geom, prec,
                                                           ws.data_ptr() if ws_bytes else None,
                                                           ws_bytes, st),
                    }
                    # ---- parity against torch CPU float64 autograd on the first `sub` images
                    pt, pl = geom.pad_t, geom.pad_l
                    pb = max((OH - 1) * s + k - H - pt, 0)
                    pr = max((OW - 1) * s + k - W - pl, 0)
                    x64 = xt[:sub].double().cpu().permute(0, 3, 1, 2).requires_grad_(True)
                    w64 = wt.double().cpu().permute(3, 2, 0, 1).requires_grad_(True)
                    y64 = TF.conv2d(TF.pad(x64, (pl, pr, pt, pb)), w64, stride=s)
                    y64.backward(gt[:sub].double().cpu().permute(0, 3, 1, 2))
                    for fn in ops.values():
                        fn()
                    torch.cuda.synchronize()
                    errs = {
                        "fprop": err(yt[:sub], y64.detach().permute(0, 2, 3, 1)),
                        "dgrad": err(dxt[:sub], x64.grad.permute(0, 2, 3, 1)),
                        "wgrad": err(dwt, w64.grad.permute(2, 3, 1, 0)),
                    }
                    flops = 2.0 * N * OH * OW * K * k * k * C
                    byts = {"fprop": 4.0 * (xt.numel() + wt.numel() + yt.numel()),
                            "dgrad": 4.0 * (gt.numel() + wt.numel() + dxt.numel()),
                            "wgrad": 4.0 * (xt.numel() + gt.numel() + wt.numel())}
                    # ---- NumPy port beside it (forward + both gradients), where it fits
                    np_ms = None
                    patch_bytes = 8.0 * N * OH * OW * k * k * C
                    if not args.no_numpy and patch_bytes <= (2 << 30) and flops <= 3e11:
                        xn = xt.cpu().numpy().astype(np.float64)
                        wn = wt.cpu().numpy().astype(np.float64)
                        gn = gt.cpu().numpy().astype(np.float64)
                        t0 = time.perf_counter()
                        orc.conv2d_grads(xn, wn, gn, "SAME", (s, s))
                        np_ms = (time.perf_counter() - t0) * 1e3
                        del xn, wn, gn
                    gpu_total = 0.0
                    for op, fn in ops.items():
                        ms = timeit(fn)
                        gpu_total += ms
                        t_tensor = flops / (tpeak * 1e12) * 1e3
                        t_hbm = byts[op] / (hbm * 1e9) * 1e3
                        bound = "tensor" if t_tensor >= t_hbm else "hbm"
                        row = {"op": f"conv_{op}", "N": N, "C": C, "K": K, "k": k, "s": s,
                               "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12,
                               "gbs": byts[op] / (ms * 1e-3) / 1e9, "bound": bound,
                               "frac": max(t_tensor, t_hbm) / ms,
                               "frac_vs_cublas_tf32": (max(flops / (cublas_tf32 * 1e12) * 1e3, t_hbm) / ms
                                                       if cublas_tf32 else None),
                               "err_vs_torch_f64": errs[op]}
                        rows.append(row)
                    if np_ms is not None:
                        rows.append({"op": "conv_fwd+bwd_numpy_port", "N": N, "C": C, "K": K, "k": k,
                                     "s": s, "ms": np_ms, "gpu_ms": gpu_total,
                                     "speedup": np_ms / gpu_total})
                    worst = max(errs.values())
                    print(f"N{N} C{C} k{k} s{s}: " + " ".join(
                        f"{r['op'][5:]} {r['ms']:.3f}ms {r['frac']:.2f}{r['bound'][0]}"
                        for r in rows[-(4 if np_ms is not None else 3):][:3]) +
                        f" | err {worst:.1e}" + (f" | numpy {np_ms:.0f} ms" if np_ms else ""),
                        flush=True)
                    assert worst < 2e-3, (N, C, k, s, errs)
                    del yt, gt, dxt, dwt, ws
            # ---- maxpool2d 2x2 stride 2 and 3x3 stride 2 on the same activation tensor
            for kp, sp_ in ((2, 2), (3, 2)):
                geomp, oshape, osize = core._pool_plan((N, H, W, C), kp, kp, sp_, sp_, "SAME")
                yt, yd = dev(oshape)
                it = torch.empty(oshape, dtype=torch.uint8, device="cuda")
                gt, gd = dev(oshape)
                dxt, dxd = dev((N, H, W, C))
                st = stream()
                fwd = lambda: F.sp_maxpool2d_fwd(xd.ptr, yd.ptr, it.data_ptr(), *geomp, st)  # noqa: E731
                bwd = lambda: F.sp_maxpool2d_bwd(gd.ptr, it.data_ptr(), dxd.ptr, *geomp, st)  # noqa: E731
                fwd()
                bwd()
                torch.cuda.synchronize()
                sub = min(N, 2)
                # independent check: torch CPU float64 max_pool2d with the reference's ZERO
                # padding (the pad value competes, sp.py:289): pad explicitly, then pool
                pt, pl, OHp, OWp = geomp[8], geomp[9], geomp[10], geomp[11]
                pb = max((OHp - 1) * sp_ + kp - H - pt, 0)
                pr = max((OWp - 1) * sp_ + kp - W - pl, 0)
                x64 = xt[:sub].double().cpu().permute(0, 3, 1, 2).requires_grad_(True)
                y64 = TF.max_pool2d(TF.pad(x64, (pl, pr, pt, pb)), kp, sp_)
                y64.backward(gt[:sub].double().cpu().permute(0, 3, 1, 2))
                e_f = err(yt[:sub], y64.detach().permute(0, 2, 3, 1))
                e_b = err(dxt[:sub], x64.grad.permute(0, 2, 3, 1))
                ins, outs = xt.numel(), yt.numel()
                for name, fn, e in (("maxpool_fwd", fwd, e_f), ("maxpool_bwd", bwd, e_b)):
                    ms = timeit(fn)
                    nb = 4.0 * ins + 5.0 * outs
                    rows.append({"op": name, "N": N, "C": C, "k": kp, "s": sp_, "ms": ms,
                                 "gbs": nb / (ms * 1e-3) / 1e9, "bound": "hbm",
                                 "frac": nb / (ms * 1e-3) / 1e9 / hbm, "err_vs_torch_f64": e})
                print(f"N{N} C{C} pool{kp}s{sp_}: fwd {rows[-2]['ms']:.3f}ms {rows[-2]['frac']:.2f}h "
                      f"bwd {rows[-1]['ms']:.3f}ms {rows[-1]['frac']:.2f}h | err {max(e_f, e_b):.1e}",
                      flush=True)
                # values and routing exact; overlapping windows ADD gradients (fp32 vs float64 sums)
                assert e_f == 0.0 and e_b <= (0.0 if kp == sp_ else 1e-6), (N, C, kp, e_f, e_b)
                del yt, it, gt, dxt
            del xt, xd
            torch.cuda.empty_cache()

    conv = [r for r in rows if r["op"].startswith("conv_") and "frac" in r]
    summary = {
        "points": len(conv),
        "worst_err_vs_torch_f64": max(r["err_vs_torch_f64"] for r in conv),
        "median_frac": float(np.median([r["frac"] for r in conv])),
        "worst": sorted(conv, key=lambda r: r["frac"])[:12],
        "seconds": time.time() - t_start,
    }
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump({"grid": {"N": Ns, "C=K": Cs, "kernel": [1, 3, 5], "stride": [1, 2],
                            "image": [H, W], "padding": "SAME"},
                   "precision": "tf32 single pass (PRECISION_TF32)",
                   "tensor_peak_tflops": tpeak, "tensor_peak_note": tpeak_note, "hbm_gbs": hbm,
                   "frac": "max(flops / tensor peak, algorithmic bytes / HBM peak) / measured time",
                   "timing": f"median of {args.reps} CUDA-event timings, L2 flushed before each",
                   "parity": "torch CPU float64 conv2d / max_pool2d autograd, full-size launches",
                   "summary": summary, "rows": rows}, fh, indent=1)
    print(json.dumps(summary)[:2000])


if __name__ == "__main__":
    main()
