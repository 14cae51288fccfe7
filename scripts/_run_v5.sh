for H in 0 1; do
export SP_PROBE_HALO=$H
echo "=== probe halo=$H"
timeout 600 python scripts/tma2_probe.py all > gpurun_out/probe_h$H.log 2>&1; grep -c PASS gpurun_out/probe_h$H.log; grep -E "FAIL|CRASH|TIMEOUT" gpurun_out/probe_h$H.log | head
done
echo "=== timing: umma1 vs tma2(im2col)"
SP_PROBE_HALO=0 SP_PROBE_B2B=1 timeout 300 python scripts/tma2_probe.py time 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['op'], 'umma1_us', round(d['umma1_ms']*1e3,2), 'tma2_us', round(d['tma2_ms']*1e3,2), round(d['tma2_tflops'],1), round(d['tma2_frac_of_tf32_peak'],3))
" | tee gpurun_out/tma2_time_v5_b2b.log
echo "=== timing: im2col vs halo"
SP_PROBE_HALO=1 SP_PROBE_B2B=1 timeout 300 python scripts/tma2_probe.py time 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    if 'wgrad' in d['op']: continue
    print(d['op'], 'im2col_us', round(d['umma1_ms']*1e3,2), 'halo_us', round(d['tma2_ms']*1e3,2), round(d['tma2_tflops'],1), round(d['tma2_frac_of_tf32_peak'],3))
" | tee gpurun_out/halo_time_v5_b2b.log
