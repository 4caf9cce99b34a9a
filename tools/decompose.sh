#!/bin/bash
# Timing decomposition of the trunk kernels on one L2-sized tile group (1044x522 input, untiled: 136k feature px).
# NESR_B200_DEBUG_FLAGS: 1 no epilogue loads/stores, 2 no MMA, 4 no TMA loads; NESR_B200_PAIRS=0 selects the single-CTA trunk kernel
out=gpurun_out/decompose.log
: > $out
for pr in 1; do
for f in 0 1 2 4 5 6 7; do
  echo -n "pairs=$pr " >> $out
  NESR_B200_PAIRS=$pr NESR_B200_DEBUG_FLAGS=$f NESR_WARMUP=2 timeout 120 python tools/quick_bench.py 522 1044 0 10 3 2>&1 | tail -1 >> $out
done
done
cat $out
