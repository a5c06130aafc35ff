#!/bin/bash
# Round-2 GPU call L (one B200): final full GPU suite, streaming LP build A/B + ncu, bench line.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 700 python -m pytest tests -m gpu -q --durations=5 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
( echo "# streaming build (keys grouped by table slice first)"; timeout 200 python tools/build_bench.py 28 lp; echo "# CCB_BUILD_DIRECT=1 (inserts in input order)"; CCB_BUILD_DIRECT=1 timeout 200 python tools/build_bench.py 28 lp ) > $O/build_bench_lp.txt 2>&1; cat $O/build_bench_lp.txt
timeout 300 ncu --set full --clock-control none -k regex:"lp_insert_ordered_kernel|lp_audit_kernel" -c 2 -f -o $O/lp_build_streaming_full \
  python tools/build_bench.py 27 lp > $O/ncu_build_lp.log 2>&1
( time timeout 600 python bench.py --no-cpu-baseline --no-chain ) > $O/bench_c4_l.json 2> $O/bench_c4_l.err; echo "bench rc=$?" >> $O/bench_c4_l.err
cut -c1-200 $O/bench_c4_l.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_l.json')); print('build_seconds', d['build_seconds'], 'ms_per_step', d['ms_per_step'])"
( timeout 200 python __graft_entry__.py smoke ) > $O/smoke.log 2>&1; tail -1 $O/smoke.log
