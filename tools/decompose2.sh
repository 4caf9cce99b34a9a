#!/bin/bash
# Dependency-chain + epilogue trace of the trunk kernel (NESR_B200_PROF=1 build), one L2-sized group.
for f in 1024 1040; do
NESR_NUM_BLOCK=2 NESR_B200_DEBUG_FLAGS=$f NESR_WARMUP=1 timeout 120 python tools/quick_bench.py 522 1044 0 10 1 > gpurun_out/trunk_trace_$f.log 2>&1
done
tail -1 gpurun_out/trunk_trace_1024.log
