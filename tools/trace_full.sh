#!/bin/bash
# chain trace (timing-experiment library, flag 1024) of dense blocks 30..35 of the full net: tools/trace_full.sh <tag>
export NESR_B200_LIB=$PWD/neural_enhanced_super_resolution_b200/libnesr_b200_prof.so
NESR_B200_DEBUG_FLAGS=1024 NESR_WARMUP=1 timeout 120 python tools/quick_bench.py 522 1044 0 10 1 > gpurun_out/${1:-trace}_522.log 2>&1
tail -1 gpurun_out/${1:-trace}_522.log
