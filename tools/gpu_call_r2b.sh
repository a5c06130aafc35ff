#!/bin/bash
# Round-2 GPU call B (one B200): full GPU suite (no -x), chain kernel variants after the bitmap / narrowing / occupancy changes,
# texture-path gather micro-benchmark, ncu capture of the default chain kernel.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 700 python -m pytest tests -m gpu -q --durations=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
rm -f $O/chain_variants.txt
for impl in cta w1 w2 w4; do
  echo "## CCB_CHAIN_IMPL=$impl" >> $O/chain_variants.txt
  CCB_CHAIN_IMPL=$impl timeout 200 python tools/chain_bench.py 4 5 20000000 2000000 chain >> $O/chain_variants.txt 2>&1
done
echo "## w1 cf=20 chain / cf=5 lp / cf=1 chain" >> $O/chain_variants.txt
timeout 200 python tools/chain_bench.py 4 20 20000000 2000000 chain >> $O/chain_variants.txt 2>&1
timeout 200 python tools/chain_bench.py 4 5 20000000 2000000 lp >> $O/chain_variants.txt 2>&1
timeout 200 python tools/chain_bench.py 4 1 20000000 2000000 chain >> $O/chain_variants.txt 2>&1
grep -E "^##|threshold full|threshold none|Error|error" $O/chain_variants.txt
timeout 120 tools/gather_bench2 > $O/gather_bench2.txt 2>&1; cat $O/gather_bench2.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_warp_kernel --launch-skip 1 -c 1 -f -o $O/chain_warp_full \
  python tools/chain_bench.py 4 5 20000000 2000000 chain > $O/ncu_chain.log 2>&1
tail -2 $O/ncu_chain.log
