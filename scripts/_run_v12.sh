#!/bin/bash
mkdir -p gpurun_out
echo "=== probe cases (correctness) im2col / halo"
timeout 600 python scripts/tma2_probe.py all 2>&1 | grep CASE
SP_PROBE_HALO=1 timeout 600 python scripts/tma2_probe.py all 2>&1 | grep CASE
echo "=== property tests"
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
export SP_PROBE_B2B=1
fmt='
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d["op"].ljust(34), "%.1f us  %.0f TF  frac %.2f" % (d["tma2_ms"]*1e3, d["tma2_tflops"], d["tma2_frac_of_tf32_peak"]))
'
echo "=== im2col"
timeout 300 python scripts/tma2_probe.py time 2>&1 | python -c "$fmt"
echo "=== halo"
SP_PROBE_HALO=1 SP_PROBE_ONLY=cnn_conv2,cnn_conv3,vgg_conv1b_64_64,vgg_conv2a_64_128,vgg_conv2b_128_128 timeout 300 python scripts/tma2_probe.py time 2>&1 | python -c "$fmt"
echo "=== im2col bn=128 on cnn"
SP_TMA2_BN=128 SP_PROBE_ONLY=cnn_conv2,cnn_conv3 timeout 300 python scripts/tma2_probe.py time 2>&1 | python -c "$fmt"
echo "=== vgg step"
timeout 300 python scripts/vgg_step.py 12 2>&1 | tail -3
echo "=== bench"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['cuda_graph'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])
print(d['top_calls_one_step_ms'])
P
tail -3 gpurun_out/bench.err
