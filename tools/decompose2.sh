#!/bin/bash
# Per-pass time stamps of the trunk kernel's dependency chain (blocks 0 and 40) and epilogue accounting, one L2-sized group.
# Needs a build with NESR_B200_PROF=1 (python -m neural_enhanced_super_resolution_b200._build); rebuild without it afterwards.
NESR_NUM_BLOCK=2 NESR_B200_DEBUG_FLAGS=1024 NESR_WARMUP=1 timeout 120 python tools/quick_bench.py 522 1044 0 10 1 > gpurun_out/trunk_trace.log 2>&1
tail -1 gpurun_out/trunk_trace.log
