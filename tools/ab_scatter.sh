#!/bin/bash
# runs on the GPU box: scatter-kernel variants (built by tools/build_variant.sh into gpurun_variants/)
mkdir -p gpurun_out; out=gpurun_out/ab_scatter.txt; : > $out
V=gpurun_variants
python tools/ab_step.py "default (2 CTAs/SM, grid 4/SM)" >> $out 2>&1
CCB_SCATTER_BLOCKS_PER_SM=2 python tools/ab_step.py "default, grid 2/SM" >> $out 2>&1
CCB_SCATTER_NO_TMA=1 python tools/ab_step.py "default, no TMA" >> $out 2>&1
CCB_LIB_PATH=$V/libccb200_mb3.so CCB_SCATTER_BLOCKS_PER_SM=3 python tools/ab_step.py "minblocks 3, grid 3/SM" >> $out 2>&1
CCB_LIB_PATH=$V/libccb200_mb3.so CCB_SCATTER_BLOCKS_PER_SM=6 python tools/ab_step.py "minblocks 3, grid 6/SM" >> $out 2>&1
CCB_LIB_PATH=$V/libccb200_mb3.so CCB_SCATTER_BLOCKS_PER_SM=3 CCB_SCATTER_NO_TMA=1 python tools/ab_step.py "minblocks 3, grid 3/SM, no TMA" >> $out 2>&1
cat $out
