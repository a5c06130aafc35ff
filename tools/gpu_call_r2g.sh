#!/bin/bash
# Round-2 GPU call G (gpurun --gpus 8): pjoin with one send slot per piece at N = 8 (4 and 8 pieces), full default bench line, host PCIe ceiling.
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29630 tools/pcie_aggregate.py > $O/pcie_aggregate_n8.txt 2>&1; grep -v "^\[" $O/pcie_aggregate_n8.txt | tail -5
for B in 8 4; do
  ( CCB_PJ_TRACE=1 timeout 300 $TR --master-port $((29631 + B)) bench.py --gpus 8 --sub-batches $B --steps 4 --no-e2e ) > $O/var_n8_b$B.json 2> $O/var_n8_b$B.err
  cut -c1-200 $O/var_n8_b$B.json; grep "pjoin timeline rank 0" $O/var_n8_b$B.err | tail -3 | head -2
done
( time timeout 500 $TR --master-port 29640 bench.py --gpus 8 ) > $O/bench_n8_default.json 2> $O/bench_n8_default.err; echo "rc=$?" >> $O/bench_n8_default.err
cut -c1-300 $O/bench_n8_default.json; tail -3 $O/bench_n8_default.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( time timeout 400 $TR4 --master-port 29641 bench.py --gpus 4 --no-e2e ) > $O/bench_n4_default.json 2> $O/bench_n4_default.err; cut -c1-200 $O/bench_n4_default.json
