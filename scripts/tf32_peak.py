"""Calibrate the dense TF32 tensor-core peak of this GPU once ("builder must
calibrate once and record").  Off the hot path: a plain library GEMM (cuBLAS through
torch.matmul with TF32 allowed) at 8192^3 and 16384 x 8192 x 8192, burst (a single timed call
after warm-up) and sustained (200 back-to-back calls).  Writes profiles/r02_tf32_peak.json, which
bench.py and the sweep scripts use as the TF32 roofline (falling back to 1/2 x the measured
bf16 figure of MEASURED_PEAKS.json when it is absent).

    python scripts/tf32_peak.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
out = {"how": "torch.matmul (cuBLAS), fp32 inputs with TF32 allowed; CUDA events", "shapes": {}}
best_burst, best_sustained = 0.0, 0.0
for m, n, k in ((8192, 8192, 8192), (16384, 8192, 8192), (4096, 4096, 16384)):
    a = torch.randn(m, k, device="cuda")
    b = torch.randn(k, n, device="cuda")
    for _ in range(5):
        c = a @ b
    torch.cuda.synchronize()
    flops = 2.0 * m * n * k
    bursts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c = a @ b
        e1.record()
        torch.cuda.synchronize()
        bursts.append(flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200
    e0.record()
    for _ in range(reps):
        c = a @ b
    e1.record()
    torch.cuda.synchronize()
    sustained = flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    out["shapes"][f"{m}x{n}x{k}"] = {"burst_tflops": max(bursts), "sustained_tflops": sustained}
    best_burst = max(best_burst, max(bursts))
    best_sustained = max(best_sustained, sustained)
    del a, b, c
out["tf32_tflops"] = best_burst
out["tf32_tflops_sustained"] = best_sustained
peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(peaks_path):
    with open(peaks_path) as fh:
        p = json.load(fh)
    out["half_bf16_burst_for_comparison"] = p.get("bf16_tflops", 0) / 2
print(json.dumps(out, indent=1))
path = os.path.join(ROOT, "gpurun_out", "r02_tf32_peak.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
with open(path, "w") as fh:
    json.dump(out, fh, indent=1)
