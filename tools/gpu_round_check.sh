#!/bin/bash
# Runs on the GPU box: GPU parity tests, the default bench, the ncu launch list of the bench
# and one full ncu capture of each dominant kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time python bench.py ) > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
echo "bench rc=$?" >> gpurun_out/bench_c4.err
if [ "$1" != "quick" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_ncu.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k regex:probe_unique_kernel --launch-skip 4 -c 1 -f -o gpurun_out/probe_unique_full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_probe_full.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k regex:partition_scatter_kernel --launch-skip 4 -c 1 -f -o gpurun_out/scatter_full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_scatter_full.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_c4.json | cut -c1-600
