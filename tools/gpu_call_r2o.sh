#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "chain" ) 2>&1 | tail -2
for lib in "" gpurun_variants/libccb200_minb8.so; do
  echo "## CCB_LIB_PATH=$lib" >> $O/chain_occ.txt
  CCB_LIB_PATH=$lib timeout 120 python tools/chain_bench.py 4 5 20000000 2000000 chain 2>&1 | grep "threshold full\|threshold none\|rror" >> $O/chain_occ.txt
done
CCB_CHAIN_IMPL=w1 timeout 120 python tools/chain_bench.py 4 5 20000000 2000000 chain 2>&1 | grep "threshold full" | sed 's/^/w1: /' >> $O/chain_occ.txt
timeout 120 python tools/chain_bench.py 4 20 20000000 2000000 chain 2>&1 | grep "threshold full" | sed 's/^/cf20: /' >> $O/chain_occ.txt
cat $O/chain_occ.txt
