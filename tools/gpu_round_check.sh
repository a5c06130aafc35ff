#!/bin/bash
# Runs on the GPU box: GPU parity tests, the default bench, the ncu launch list of the bench and one full ncu capture of a
# dominant kernel (arg 1: regex of its name, default partition_scatter_kernel; "none" skips it).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
K=${1:-partition_scatter_kernel}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 200 python bench.py ) > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
echo "bench rc=$?" >> gpurun_out/bench_c4.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_ncu.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
if [ "$K" != "none" ]; then
timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 4 -c 1 -f -o gpurun_out/kernel_full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_kernel_full.log 2>&1
fi
timeout 90 python tools/c1_bench.py > gpurun_out/c1_bench.txt 2>&1
tail -3 gpurun_out/pytest_gpu.log; cut -c1-300 gpurun_out/bench_c4.json; cat gpurun_out/c1_bench.txt
