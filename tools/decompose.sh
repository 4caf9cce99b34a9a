#!/bin/bash
# TMA-only timing of the trunk kernels on one L2-sized tile group (flags 3 = no epilogue, no MMA; 7 = nothing).
out=gpurun_out/decompose.log
: > $out
for pr in 1 0; do
for f in 3 7; do
  echo -n "pairs=$pr " >> $out
  NESR_B200_PAIRS=$pr NESR_B200_DEBUG_FLAGS=$f NESR_WARMUP=2 timeout 120 python tools/quick_bench.py 522 1044 0 10 3 2>&1 | tail -1 >> $out
done
done
cat $out
