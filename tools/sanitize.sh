#!/bin/bash
# Runs on the GPU box: compute-sanitizer memcheck + racecheck over a small, fast subset of the GPU parity tests
# (SURVEY section 4 test plan); the logs of the last run are kept under profiles/ (r2_sanitizer_*.log).
# The subset covers every kernel family once: table builds, chunk protocol, batch probe (lean + generic + payload,
# direct + partitioned), partition / scatter (TMA ring), fused chain, compactor.
mkdir -p gpurun_out
SEL='hash_bit_exact or (lp_build_equals_reference_layout and 1024) or (probe_batch_matches_oracle and 2000-4) or partition_single_regions or (probe_batch_payload_matches_oracle and 2000-4) or (partition_kernels and 12) or chain_execute_edge_cases or chain_execute_materialized_tuples or probe_batch_skewed_keys'
for tool in memcheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 97 --target-processes all \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_payload.py tests/test_gpu_peer_exchange.py -x -q -k "$SEL" \
    > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
