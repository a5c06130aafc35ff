#!/bin/bash
# Round-2 GPU call A (one B200): GPU test suite, smoke, A/B of the new kernels, chain variants, sanitizers, ncu captures.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
( timeout 200 python __graft_entry__.py smoke ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
# probe kernel A/B: deferred tail (default) vs the round-1 inline tail
( timeout 150 python tools/ab_step.py "deferred tail (probe_unique_lp_kernel)"; CCB_LEAN_INLINE_TAIL=1 timeout 150 python tools/ab_step.py "inline tail (round-1 probe_unique_kernel)" ) > $O/ab_probe_tail.txt 2>&1
cat $O/ab_probe_tail.txt
# chain kernel variants on C3
for impl in cta w1 w2 w4; do
  echo "## CCB_CHAIN_IMPL=$impl" >> $O/chain_variants.txt
  CCB_CHAIN_IMPL=$impl timeout 200 python tools/chain_bench.py 4 5 20000000 2000000 chain >> $O/chain_variants.txt 2>&1
done
CCB_CHAIN_IMPL=w4 timeout 200 python tools/chain_bench.py 4 20 20000000 2000000 chain >> $O/chain_variants.txt 2>&1
CCB_CHAIN_IMPL=w4 timeout 200 python tools/chain_bench.py 4 5 20000000 2000000 lp >> $O/chain_variants.txt 2>&1
grep -E "^##|threshold full|threshold none" $O/chain_variants.txt
# build side
timeout 300 python tools/build_bench.py 28 > $O/build_bench.txt 2>&1; cat $O/build_bench.txt
# ncu: chain kernel (C3, full compaction launch), build kernels at 2^27 keys
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_warp_kernel --launch-skip 1 -c 1 -f -o $O/chain_warp_full \
  python tools/chain_bench.py 4 5 20000000 2000000 chain > $O/ncu_chain.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"lp_insert_ordered_kernel|chain_count_kernel|chain_scatter_kernel|chain_finalize_kernel" -c 4 -f -o $O/build_kernels_full \
  python tools/build_bench.py 27 > $O/ncu_build.log 2>&1
# sanitizers
bash tools/sanitize.sh > $O/sanitize_summary.txt 2>&1; cat $O/sanitize_summary.txt
# bench (no CPU baseline in this call)
( time timeout 400 python bench.py --no-cpu-baseline ) > $O/bench_c4.json 2> $O/bench_c4.err; echo "bench rc=$?" >> $O/bench_c4.err
cut -c1-600 $O/bench_c4.json; tail -3 $O/bench_c4.err
