export SP_PROBE_HALO=1
timeout 600 python scripts/tma2_probe.py all > gpurun_out/probe_h1.log 2>&1; echo "probe halo pass: $(grep -c PASS gpurun_out/probe_h1.log)"; grep -E "FAIL|CRASH|TIMEOUT" gpurun_out/probe_h1.log | head
SP_PROBE_B2B=1 SP_PROBE_ONLY=cnn_conv2,cnn_conv3,vgg_conv1b_64_64,vgg_conv2a_64_128,vgg_conv2b_128_128,vgg_conv3a_128_256,vgg_conv3b_256_256 timeout 300 python scripts/tma2_probe.py time 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    if 'wgrad' in d['op']: continue
    print(d['op'], 'im2col_us', round(d['umma1_ms']*1e3,2), 'halo_us', round(d['tma2_ms']*1e3,2), round(d['tma2_tflops'],1), round(d['tma2_frac_of_tf32_peak'],3))
" | tee gpurun_out/halo_time_v8_b2b.log
