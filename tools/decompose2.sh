#!/bin/bash
# Dependency-chain trace of the trunk kernel (NESR_B200_PROF=1 build), one L2-sized group.
NESR_NUM_BLOCK=2 NESR_B200_DEBUG_FLAGS=1024 NESR_WARMUP=1 timeout 120 python tools/quick_bench.py 522 1044 0 10 1 > gpurun_out/trunk_trace.log 2>&1
tail -1 gpurun_out/trunk_trace.log
