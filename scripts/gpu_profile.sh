#!/bin/bash
# Tests + bench + op sweep + ncu launch list + one ncu --set full capture of the tcgen05 kernel.
mkdir -p gpurun_out
echo "=== full gpu suite" | tee gpurun_out/pytest_gpu.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider >> gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
echo "=== bench"
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "=== op sweep"
timeout 900 python scripts/op_sweep.py --reps 10 > gpurun_out/op_sweep.log 2>&1; tail -60 gpurun_out/op_sweep.log
if [ "$1" == "ncu" ]; then
echo "=== ncu launch list"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 260 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
echo "=== ncu full capture (tcgen05 kernel, VGG conv)"
timeout 300 python scripts/op_sweep.py --quick --reps 2 > gpurun_out/ncu_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm_umma -s 30 -c 4 \
   -o gpurun_out/prof_umma -f python scripts/op_sweep.py --quick --reps 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
