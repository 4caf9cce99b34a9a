#!/bin/bash
# A/B of library builds inside ONE gpurun call (box-to-box variance is +-5 %): tools/ab.sh <tag> lib1.so lib2.so ...
tag=$1; shift
out=gpurun_out/ab_$tag.log
: > $out
for rep in 1 2; do
  for lib in "$@"; do
    for cfg in "${CFGS[@]:-522 1044 0 10;1080 1920 512 10}"; do :; done
    for cfg in "522 1044 0 10" "1080 1920 512 10"; do
      echo -n "$lib $cfg : " >> $out
      NESR_B200_LIB=$PWD/neural_enhanced_super_resolution_b200/$lib NESR_WARMUP=3 timeout 120 python tools/quick_bench.py $cfg 5 2>&1 | tail -1 >> $out
    done
  done
done
cat $out
