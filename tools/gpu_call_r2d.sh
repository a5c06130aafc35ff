#!/bin/bash
# Round-2 GPU call D (gpurun --gpus N): pjoin after the copy-split / two-group probe changes: pjoin tests (sliced small tables),
# quick parity with the fused path forced, bench cabi with 4 and 8 pieces + timelines.
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 600 python -m pytest tests/test_gpu_multirank.py -q -k "pjoin" ) > $O/pytest_pjoin_n$N.log 2>&1; tail -4 $O/pytest_pjoin_n$N.log
( time timeout 600 $TR --master-port 29612 tests/multirank_parity.py --quick --partitioned-probe --report $O/multirank_parity_p${N}_partitioned_probe.txt ) > $O/parity_pp_p$N.out 2>&1
grep -c "^PASS" $O/multirank_parity_p${N}_partitioned_probe.txt; grep "^FAIL\|^#" $O/multirank_parity_p${N}_partitioned_probe.txt | head; tail -3 $O/parity_pp_p$N.out
for B in 4 8; do
  CCB_PJ_TRACE=1 timeout 300 $TR --master-port 29615 bench.py --gpus $N --exchange cabi --sub-batches $B --steps 3 --no-e2e 2>&1 | grep "pjoin timeline rank 0" | tail -3 | tee $O/pj_trace_n${N}_b$B.txt
  ( time timeout 600 $TR --master-port 29613 bench.py --gpus $N --exchange cabi --sub-batches $B ) > $O/bench_n${N}_cabi_b$B.json 2> $O/bench_n${N}_cabi_b$B.err; echo "rc=$?" >> $O/bench_n${N}_cabi_b$B.err
  cut -c1-220 $O/bench_n${N}_cabi_b$B.json; tail -2 $O/bench_n${N}_cabi_b$B.err
done
( time timeout 600 $TR --master-port 29616 bench.py --gpus $N --exchange cabi --no-pipeline --no-e2e ) > $O/bench_n${N}_cabi_nopipe.json 2> $O/bench_n${N}_cabi_nopipe.err; cut -c1-220 $O/bench_n${N}_cabi_nopipe.json
L=0; n=$N; while [ $n -gt 1 ]; do n=$((n / 2)); L=$((L + 1)); done
timeout 300 chunk-compaction-in-vectorized-execution-simd_b200/host/pjoin_main --gpus $N --log2-build $((27 + L)) --log2-probe $((30 + L)) --steps 3 --pipeline 1 --sub-batches 4 > $O/pjoin_main_n$N.json 2> $O/pjoin_main_n$N.err; cat $O/pjoin_main_n$N.json; tail -2 $O/pjoin_main_n$N.err
