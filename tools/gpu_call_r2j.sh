#!/bin/bash
# Round-2 GPU call J (one B200): the round's final single-GPU record -- full GPU suite, smoke, default bench (all legs), reference
# arm, ncu launch list of the bench command, ncu --set full of the chain kernel and of the table-build kernels.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
( time timeout 700 python -m pytest tests -m gpu -q --durations=5 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
( timeout 200 python __graft_entry__.py smoke ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
( time timeout 600 python bench.py ) > $O/bench_c4.json 2> $O/bench_c4.err; echo "bench rc=$?" >> $O/bench_c4.err
cut -c1-400 $O/bench_c4.json; tail -4 $O/bench_c4.err
( time timeout 400 python bench.py --impl reference --steps 5 --warmup 2 ) > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-400 $O/bench_reference.json; tail -3 $O/bench_reference.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launches_ncu.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_warp_kernel --launch-skip 1 -c 1 -f -o $O/chain_warp_full \
  python tools/chain_bench.py 4 5 20000000 2000000 chain > $O/ncu_chain.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"chain_count_kernel|chain_scatter_kernel|chain_finalize_kernel|occupancy_kernel" -c 4 -f -o $O/chain_build_full \
  python tools/build_bench.py 27 chain > $O/ncu_build_chain.log 2>&1
timeout 120 python tools/chain_bench.py 4 5 20000000 2000000 chain > $O/chain_final.txt 2>&1; grep threshold $O/chain_final.txt
