#!/bin/bash
# Round-2 GPU call I (gpurun --gpus 2): parity gate incl. the host-buffer paths, default bench line at N = 2.
mkdir -p gpurun_out
O=gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29611 tests/multirank_parity.py --report $O/multirank_parity_p$N.txt ) > $O/parity_p$N.out 2>&1; echo "parity rc=$?" >> $O/parity_p$N.out
grep -c "^PASS" $O/multirank_parity_p$N.txt; grep "^FAIL\|^#" $O/multirank_parity_p$N.txt | head -20; tail -4 $O/parity_p$N.out
( time timeout 600 $TR --master-port 29613 bench.py --gpus $N ) > $O/bench_n2_default.json 2> $O/bench_n2_default.err; echo "rc=$?" >> $O/bench_n2_default.err
cat $O/bench_n2_default.json | cut -c1-1500; tail -3 $O/bench_n2_default.err
