#!/bin/bash
# Round-2 GPU call K (one B200): re-check after the scatter-kernel template fix -- failed tests, C4 step A/B, bench line, launch list.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q -k "telemetry or pjoin or partition or skewed or probe_batch_large or peer_exchange or copy_exchange" ) > $O/pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_k.log
tail -6 $O/pytest_gpu_k.log
timeout 150 python tools/ab_step.py "final (FUSED as a template parameter)" > $O/ab_scatter_template.txt 2>&1; cat $O/ab_scatter_template.txt
( time timeout 600 python bench.py --no-cpu-baseline ) > $O/bench_c4.json 2> $O/bench_c4.err; echo "bench rc=$?" >> $O/bench_c4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c4.json"))
r = d["roofline"]
print(d["ms_per_step"], d["value"] / 1e9, r["frac"], r["hw_frac"], [(k["kernel"][:24], round(k["ms"], 2)) for k in r["kernels"]])
for x in d["chain"]["runs"]:
    print(x["chunk_factor"], x["policy"][:40], round(x["ms"], 3), x.get("device_ms"))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launches_ncu.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-chain > $O/ncu_launches.log 2>&1
