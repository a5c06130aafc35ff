#!/bin/bash
# Round-2 GPU call Q (gpurun --gpus 8): SM copy share of the C-ABI join at N = 8 (the step is copy-bound there).
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for pct in $@; do
  ( CCB_PJ_SM_COPY_PCT=$pct timeout 150 $TR --master-port $((29700 + pct)) bench.py --gpus 8 --steps 6 --no-e2e ) > $O/sm_n8_$pct.json 2> $O/sm_n8_$pct.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/sm_n8_$pct.json").read().strip().splitlines()[-1])
    print(f"N = 8, SM copy share $pct %: {d['ms_per_step']:.2f} ms/step  {d['value'] / 1e9:.1f} G tuples/s  checks {d['checks']}")
except Exception as e:
    print("SM copy share $pct %: FAILED", e)
PY
done
