#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/ab_cursor.txt; : > $out
(timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer_exchange.py -x -q -k "partition or probe_batch or scatter or exchange or stream") > gpurun_out/pytest_part.log 2>&1; tail -2 gpurun_out/pytest_part.log >> $out
CCB_CURSOR_STRIDE=1 python tools/ab_step.py "cursor stride 1 (dense)" >> $out 2>&1
CCB_CURSOR_STRIDE=4 python tools/ab_step.py "cursor stride 4 (one per 32 B sector)" >> $out 2>&1
CCB_CURSOR_STRIDE=16 python tools/ab_step.py "cursor stride 16 (one per 128 B line)" >> $out 2>&1
CCB_CURSOR_STRIDE=64 python tools/ab_step.py "cursor stride 64 (512 B apart)" >> $out 2>&1
cat $out
