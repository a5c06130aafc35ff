#!/bin/bash
# scatter-kernel ablation (measurement builds, results are wrong by design): which part of the tile loop costs the time?
mkdir -p gpurun_out; out=gpurun_out/ab_ablate.txt; : > $out
V=gpurun_variants
python tools/ab_step.py "full kernel" >> $out 2>&1
AB_NOCHECK=1 CCB_LIB_PATH=$V/libccb200_abl1.so python tools/ab_step.py "ablate 1: no output stores" >> $out 2>&1
AB_NOCHECK=1 CCB_LIB_PATH=$V/libccb200_abl2.so python tools/ab_step.py "ablate 2: no global range atomics" >> $out 2>&1
AB_NOCHECK=1 CCB_LIB_PATH=$V/libccb200_abl3.so python tools/ab_step.py "ablate 3: no smem ranking atomics" >> $out 2>&1
AB_NOCHECK=1 CCB_LIB_PATH=$V/libccb200_abl4.so python tools/ab_step.py "ablate 4: no smem sort (identity slots)" >> $out 2>&1
cat $out
