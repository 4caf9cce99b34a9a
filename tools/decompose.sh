#!/bin/bash
# Epilogue decomposition of the trunk kernel on one L2-sized tile group.
# debug flags: 1 no epilogue loads/stores, 2 no MMA, 4 no TMA loads, 16 no 16-bit activation stores, 4096 no fp32 trunk stores
out=gpurun_out/decompose.log
: > $out
for f in 6 22 4118 7 16 0; do
  NESR_CONV_IMPL=0 NESR_B200_DEBUG_FLAGS=$f NESR_WARMUP=2 timeout 120 python tools/quick_bench.py 522 1044 0 10 3 2>&1 | tail -1 >> $out
done
cat $out
