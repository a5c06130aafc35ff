#!/bin/bash
# Round-2 GPU call H (gpurun --gpus 8): final N = 8 numbers (two batches of send slots), 8 pieces as a variant, C++ driver, N = 4.
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( CCB_PJ_TRACE=1 timeout 300 $TR --master-port 29639 bench.py --gpus 8 --sub-batches 8 --steps 4 --no-e2e ) > $O/var_n8_b8.json 2> $O/var_n8_b8.err
cut -c1-200 $O/var_n8_b8.json; grep "pjoin timeline rank 0" $O/var_n8_b8.err | tail -3 | head -2
( CCB_PJ_TRACE=1 timeout 300 $TR --master-port 29638 bench.py --gpus 8 --steps 4 --no-e2e ) > $O/var_n8_b4.json 2> $O/var_n8_b4.err
grep "pjoin timeline rank 0" $O/var_n8_b4.err | tail -3 | head -2
( time timeout 500 $TR --master-port 29640 bench.py --gpus 8 ) > $O/bench_n8_default.json 2> $O/bench_n8_default.err; echo "rc=$?" >> $O/bench_n8_default.err
cut -c1-300 $O/bench_n8_default.json; tail -3 $O/bench_n8_default.err
timeout 300 chunk-compaction-in-vectorized-execution-simd_b200/host/pjoin_main --gpus 8 --log2-build 30 --log2-probe 33 --steps 3 --pipeline 1 --sub-batches 4 > $O/pjoin_main_n8.json 2> $O/pjoin_main_n8.err; cat $O/pjoin_main_n8.json; tail -2 $O/pjoin_main_n8.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( time timeout 400 $TR4 --master-port 29641 bench.py --gpus 4 --no-e2e ) > $O/bench_n4_default.json 2> $O/bench_n4_default.err; cut -c1-200 $O/bench_n4_default.json
