#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( echo "# streaming build (keys grouped by table slice first, resident CTAs only)"; timeout 200 python tools/build_bench.py 28 lp; echo "# CCB_BUILD_DIRECT=1 (inserts in input order)"; CCB_BUILD_DIRECT=1 timeout 200 python tools/build_bench.py 28 lp ) > $O/build_bench_lp.txt 2>&1; cat $O/build_bench_lp.txt
timeout 300 ncu --set full --clock-control none -k regex:"lp_insert_ordered_kernel|lp_audit_kernel" -c 2 -f -o $O/lp_build_streaming_full \
  python tools/build_bench.py 27 lp > $O/ncu_build_lp.log 2>&1
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lp_streaming or lp_build or probe_batch_large" ) 2>&1 | tail -2
