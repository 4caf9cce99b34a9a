#!/bin/bash
# Timing decomposition of the conv kernels (needs a NESR_B200_PROF=1 build for the role/pass traces).
# debug flags: 1 no epilogue loads/stores, 2 no MMA, 4 no TMA loads, 8 no TMEM re-zero, 32 role trace, 1024 pass trace
out=gpurun_out/decompose.log
: > $out
for f in 0 1 2 4 3 5 6 7; do
  NESR_B200_DEBUG_FLAGS=$f NESR_WARMUP=2 timeout 120 python tools/quick_bench.py 1080 1920 512 10 3 2>&1 | tail -1 >> $out
done
NESR_NUM_BLOCK=2 NESR_B200_DEBUG_FLAGS=1056 NESR_WARMUP=1 timeout 120 python tools/quick_bench.py 1080 1920 512 10 1 > gpurun_out/body_trace.log 2>&1
cat $out
