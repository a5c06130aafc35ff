#!/bin/bash
# Round-2 GPU call F (gpurun --gpus N): pjoin variants -- copy streams, direct NVLink stores of the scatter kernel, pieces.
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29620
run() {  # tag, env..., -- bench args
  tag=$1; shift
  port=$((port + 1))
  ( env "$@" CCB_PJ_TRACE=1 timeout 300 $TR --master-port $port bench.py --gpus $N --exchange cabi --steps 4 --no-e2e $EXTRA ) > $O/var_n${N}_$tag.json 2> $O/var_n${N}_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/var_n${N}_$tag.json").read().strip().splitlines()[-1])
    print(f"$tag: {d['ms_per_step']:.2f} ms/step  {d['value'] / 1e9:.1f} G tuples/s  (sub_batches {d['config']['sub_batches']})")
except Exception as e:
    print("$tag: FAILED", e)
PY
  grep "pjoin timeline rank 0" $O/var_n${N}_$tag.err | tail -1
}
EXTRA="--sub-batches ${2:-4}"
run cs4_direct0 CCB_PJ_COPY_STREAMS=4 CCB_PJ_DIRECT=0
run cs8_direct0 CCB_PJ_COPY_STREAMS=8 CCB_PJ_DIRECT=0
run cs4_direct${3:-1} CCB_PJ_COPY_STREAMS=4 CCB_PJ_DIRECT=${3:-1}
if [ -n "$4" ]; then run cs4_direct$4 CCB_PJ_COPY_STREAMS=4 CCB_PJ_DIRECT=$4; fi
