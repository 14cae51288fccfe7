#!/bin/bash
mkdir -p gpurun_out
export SP_PROBE_B2B=1 SP_PROBE_ONLY=cnn_conv2,cnn_conv3,vgg_conv1b_64_64,vgg_conv2a_64_128,vgg_conv2b_128_128,vgg_conv3a_128_256
echo "=== stages=3"
SP_TMA2_STAGES=3 timeout 300 python scripts/tma2_probe.py time 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['op'].ljust(34), 'tma2 %.1f us  %.0f TF  frac %.2f' % (d['tma2_ms']*1e3, d['tma2_tflops'], d['tma2_frac_of_tf32_peak']))
"
echo "=== stages=default"
timeout 300 python scripts/tma2_probe.py time 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['op'].ljust(34), 'tma2 %.1f us  %.0f TF  frac %.2f' % (d['tma2_ms']*1e3, d['tma2_tflops'], d['tma2_frac_of_tf32_peak']))
"
echo "=== probe cases (correctness)"
timeout 600 python scripts/tma2_probe.py all 2>&1 | grep CASE
echo "=== vgg step"
timeout 300 python scripts/vgg_step.py 12 2>&1 | tail -3
