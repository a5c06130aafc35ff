#!/bin/bash
# Round-2 GPU call C (run with gpurun --gpus 2): multi-rank parity gate, the C++ pjoin driver, bench at N = 2 (ce and cabi).
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29611 tests/multirank_parity.py --report $O/multirank_parity_p$N.txt ) > $O/parity_p$N.out 2>&1; echo "parity rc=$?" >> $O/parity_p$N.out
grep -c "^PASS" $O/multirank_parity_p$N.txt; grep "^FAIL\|^#" $O/multirank_parity_p$N.txt | head -20; tail -5 $O/parity_p$N.out
( time timeout 600 $TR --master-port 29612 tests/multirank_parity.py --quick --partitioned-probe --report $O/multirank_parity_p${N}_partitioned_probe.txt ) > $O/parity_pp_p$N.out 2>&1
grep -c "^PASS" $O/multirank_parity_p${N}_partitioned_probe.txt; grep "^FAIL\|^#" $O/multirank_parity_p${N}_partitioned_probe.txt | head; tail -3 $O/parity_pp_p$N.out
( time timeout 600 python -m pytest tests/test_gpu_multirank.py -q -k "pjoin" ) > $O/pytest_pjoin_n$N.log 2>&1; tail -6 $O/pytest_pjoin_n$N.log
for ex in ce cabi; do
  ( time timeout 600 $TR --master-port 29613 bench.py --gpus $N --exchange $ex ) > $O/bench_n${N}_$ex.json 2> $O/bench_n${N}_$ex.err; echo "rc=$?" >> $O/bench_n${N}_$ex.err
  cut -c1-300 $O/bench_n${N}_$ex.json; tail -4 $O/bench_n${N}_$ex.err
done
CCB_CE_TRACE=1 timeout 300 $TR --master-port 29614 bench.py --gpus $N --exchange ce --ce-probe stream --steps 2 --no-e2e > $O/ce_trace_n$N.txt 2>&1
grep "ce timeline" $O/ce_trace_n$N.txt | tail -2
CCB_PJ_TRACE=1 timeout 300 $TR --master-port 29615 bench.py --gpus $N --exchange cabi --steps 2 --no-e2e 2>&1 | grep "pjoin timeline rank 0" | tail -2 | tee $O/pj_trace_n$N.txt
L=0; n=$N; while [ $n -gt 1 ]; do n=$((n / 2)); L=$((L + 1)); done
timeout 300 chunk-compaction-in-vectorized-execution-simd_b200/host/pjoin_main --gpus $N --log2-build $((27 + L)) --log2-probe $((30 + L)) --steps 3 > $O/pjoin_main_n$N.json 2> $O/pjoin_main_n$N.err; cat $O/pjoin_main_n$N.json; tail -2 $O/pjoin_main_n$N.err
