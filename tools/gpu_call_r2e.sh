#!/bin/bash
# Round-2 GPU call E (gpurun --gpus 8): parity gate at P = 8 and P = 4, C++ pjoin driver tests, bench at N = 8 / 4 (cabi, ce).
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/topo_n8.txt 2>&1
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  ( time timeout 600 $TR --master-port 29611 tests/multirank_parity.py --report $O/multirank_parity_p$N.txt ) > $O/parity_p$N.out 2>&1; echo "parity rc=$?" >> $O/parity_p$N.out
  grep -c "^PASS" $O/multirank_parity_p$N.txt; grep "^FAIL\|^#" $O/multirank_parity_p$N.txt | head -8; tail -3 $O/parity_p$N.out
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 300 $TR --master-port 29612 tests/multirank_parity.py --quick --partitioned-probe --report $O/multirank_parity_p8_partitioned_probe.txt ) > $O/parity_pp_p8.out 2>&1
grep -c "^PASS" $O/multirank_parity_p8_partitioned_probe.txt; grep "^FAIL\|^#" $O/multirank_parity_p8_partitioned_probe.txt | head -4
( time timeout 600 python -m pytest tests/test_gpu_multirank.py -q -k "pjoin" ) > $O/pytest_pjoin_n8.log 2>&1; tail -4 $O/pytest_pjoin_n8.log
( time timeout 400 $TR --master-port 29613 bench.py --gpus 8 --exchange cabi ) > $O/bench_n8_cabi.json 2> $O/bench_n8_cabi.err; echo "rc=$?" >> $O/bench_n8_cabi.err
cut -c1-240 $O/bench_n8_cabi.json; tail -3 $O/bench_n8_cabi.err
CCB_PJ_TRACE=1 timeout 300 $TR --master-port 29615 bench.py --gpus 8 --exchange cabi --steps 3 --no-e2e 2>&1 | grep "pjoin timeline rank 0" | tail -3 | tee $O/pj_trace_n8.txt
( time timeout 400 $TR --master-port 29614 bench.py --gpus 8 --exchange cabi --sub-batches 4 --no-e2e ) > $O/bench_n8_cabi_b4.json 2> $O/bench_n8_cabi_b4.err; cut -c1-240 $O/bench_n8_cabi_b4.json
( time timeout 400 $TR --master-port 29616 bench.py --gpus 8 --exchange cabi --no-pipeline --no-e2e ) > $O/bench_n8_cabi_nopipe.json 2> $O/bench_n8_cabi_nopipe.err; cut -c1-240 $O/bench_n8_cabi_nopipe.json
( time timeout 400 $TR --master-port 29617 bench.py --gpus 8 --exchange ce ) > $O/bench_n8_ce.json 2> $O/bench_n8_ce.err; echo "rc=$?" >> $O/bench_n8_ce.err
cut -c1-240 $O/bench_n8_ce.json; tail -3 $O/bench_n8_ce.err
timeout 300 chunk-compaction-in-vectorized-execution-simd_b200/host/pjoin_main --gpus 8 --log2-build 30 --log2-probe 33 --steps 3 --pipeline 1 --sub-batches 8 > $O/pjoin_main_n8.json 2> $O/pjoin_main_n8.err; cat $O/pjoin_main_n8.json; tail -2 $O/pjoin_main_n8.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( time timeout 400 $TR4 --master-port 29618 bench.py --gpus 4 --exchange cabi --no-e2e ) > $O/bench_n4_cabi.json 2> $O/bench_n4_cabi.err; cut -c1-240 $O/bench_n4_cabi.json
