#!/bin/bash
# Round-2 GPU call N (gpurun --gpus N): the SM-assisted copy share of the C-ABI join -- correctness (parity gate with the fused path
# forced and a 20 % SM share, C++ driver tests) and its effect on the step.
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( CCB_PJ_SM_COPY_PCT=20 timeout 400 $TR --master-port 29612 tests/multirank_parity.py --quick --partitioned-probe --report $O/multirank_parity_p${N}_smcopy.txt ) > $O/parity_sm_p$N.out 2>&1
grep -c "^PASS" $O/multirank_parity_p${N}_smcopy.txt; grep "^FAIL\|^#" $O/multirank_parity_p${N}_smcopy.txt | head; tail -2 $O/parity_sm_p$N.out
( CCB_PJ_SM_COPY_PCT=30 timeout 300 python -m pytest tests/test_gpu_multirank.py -q -k "pjoin" ) 2>&1 | tail -2
for pct in ${2:-0 10 20}; do
  ( CCB_PJ_SM_COPY_PCT=$pct CCB_PJ_TRACE=1 timeout 300 $TR --master-port $((29640 + pct)) bench.py --gpus $N --steps 4 --no-e2e ) > $O/sm_n${N}_$pct.json 2> $O/sm_n${N}_$pct.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/sm_n${N}_$pct.json").read().strip().splitlines()[-1])
    print(f"SM copy share $pct %: {d['ms_per_step']:.2f} ms/step  {d['value'] / 1e9:.1f} G tuples/s")
except Exception as e:
    print("SM copy share $pct %: FAILED", e)
PY
  grep "pjoin timeline rank 0" $O/sm_n${N}_$pct.err | tail -3 | head -1
done
