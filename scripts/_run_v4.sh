timeout 600 python scripts/tma2_probe.py all > gpurun_out/tma2_probe.log 2>&1; grep -c PASS gpurun_out/tma2_probe.log; grep -E "FAIL|CRASH|TIMEOUT" gpurun_out/tma2_probe.log
timeout 300 python scripts/tma2_trace.py cnn_conv3 vgg_conv3b 2>&1 | cut -c1-700 | tee gpurun_out/tma2_trace.log
SP_PROBE_B2B=1 timeout 300 python scripts/tma2_probe.py time 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['op'], 'umma1_us', round(d['umma1_ms']*1e3,2), 'tma2_us', round(d['tma2_ms']*1e3,2), round(d['tma2_tflops'],1), round(d['tma2_frac_of_tf32_peak'],3))
" | tee gpurun_out/tma2_time_v4_b2b.log
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv2d" -p no:cacheprovider 2>&1 | tail -5
