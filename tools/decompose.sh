#!/bin/bash
# Timing decomposition of the trunk kernels on one L2-sized tile group (1044x522 input, untiled: 136k feature px).
# NESR_B200_DEBUG_FLAGS (results are wrong when set): 1 no epilogue loads/stores, 2 no MMA, 4 no TMA loads, 16 no 16-bit
# activation stores, 4096 no fp32 trunk stores, 8192 no proxy fences, 16384 no dependency polling, 32768 no publish.
# NESR_B200_PAIRS=1 selects the CTA-pair kernel (conv3x3_trunk2.cu), NESR_CONV_IMPL=4 the whole-frame kernel (conv3x3_body.cu).
# Results of r1: profiles/r1_trunk_experiments.txt
out=gpurun_out/decompose.log
: > $out
for pr in 0 1; do
for f in 0 1 2 4 3 5 6 7 57344; do
  echo -n "pairs=$pr " >> $out
  NESR_B200_PAIRS=$pr NESR_B200_DEBUG_FLAGS=$f NESR_WARMUP=2 timeout 120 python tools/quick_bench.py 522 1044 0 10 3 2>&1 | tail -1 >> $out
done
done
cat $out
