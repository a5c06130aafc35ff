#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/ab_lean_blocks.txt
timeout 150 python tools/ab_step.py "lean probe kernel: 4 CTAs/SM, 64 regs (default)" >> $O/ab_lean_blocks.txt 2>&1
CCB_LIB_PATH=gpurun_variants/libccb200_lean3.so timeout 150 python tools/ab_step.py "lean probe kernel: 3 CTAs/SM, 80 regs" >> $O/ab_lean_blocks.txt 2>&1
CCB_LIB_PATH=gpurun_variants/libccb200_lean5.so timeout 150 python tools/ab_step.py "lean probe kernel: 5 CTAs/SM, 48 regs (12 B spills)" >> $O/ab_lean_blocks.txt 2>&1
CCB_SCATTER_NO_TMA=1 timeout 150 python tools/ab_step.py "default probe, scatter without the TMA ring" >> $O/ab_lean_blocks.txt 2>&1
cat $O/ab_lean_blocks.txt
