#!/bin/bash
# Role / pass traces of the trunk kernel under the timing-experiment flags (NESR_B200_PROF=1 build).
for f in 7 5 3 0; do
  NESR_NUM_BLOCK=1 NESR_B200_DEBUG_FLAGS=$((f + 32 + 1024)) NESR_WARMUP=1 timeout 120 python tools/quick_bench.py 1080 1920 512 10 1 > gpurun_out/trace_f$f.log 2>&1
done
tail -2 gpurun_out/trace_f7.log
