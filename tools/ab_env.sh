#!/bin/bash
# A/B of environment switches of the product library inside ONE gpurun call: tools/ab_env.sh <tag> "VAR=val ..." "VAR=val ..." ...
tag=$1; shift
out=gpurun_out/abenv_$tag.log
: > $out
for rep in 1 2; do
  for envs in "$@"; do
    for cfg in ${CFGS:-"1080 1920 512 10"}; do :; done
    echo -n "[$envs] : " >> $out
    env $envs NESR_WARMUP=3 timeout 120 python tools/quick_bench.py ${CFG:-1080 1920 512 10} 5 2>&1 | tail -1 >> $out
  done
done
cat $out
