#!/bin/bash
# Is the TMA-only time of the trunk kernel bandwidth or issue bound?  flags 3 = TMA only; +65536 = half the bytes per slab
out=gpurun_out/decompose.log
: > $out
for f in 3 65539 7; do
  NESR_B200_DEBUG_FLAGS=$f NESR_WARMUP=2 timeout 120 python tools/quick_bench.py 522 1044 0 10 3 2>&1 | tail -1 >> $out
done
cat $out
