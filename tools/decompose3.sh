#!/bin/bash
# Round-2 timing experiments on the row-granular trunk kernel (timing-experiment library: python -m ..._build --prof).
#   65536 proxy fence in every writer | 131072 publish whole passes only | 262144 wait for whole passes | 524288 consumer-side proxy fence too
#   1 no epilogue stores | 2 no MMA | 4 no TMA | 16384 no dependency waits (results are wrong with any of these)
export NESR_B200_LIB=$PWD/neural_enhanced_super_resolution_b200/libnesr_b200_prof.so
out=gpurun_out/${1:-decompose3}.log
: > $out
for f in ${FLAGS:-0 65536 131072 262144 393216 458752 8192 16384}; do
  NESR_B200_DEBUG_FLAGS=$f NESR_WARMUP=2 timeout 120 python tools/quick_bench.py 522 1044 0 10 3 2>&1 | tail -1 >> $out
done
for f in ${FLAGS2:-0 65536 393216}; do
  NESR_B200_DEBUG_FLAGS=$f NESR_WARMUP=2 timeout 120 python tools/quick_bench.py 1080 1920 512 10 3 2>&1 | tail -1 >> $out
done
NESR_NUM_BLOCK=2 NESR_B200_DEBUG_FLAGS=1024 NESR_WARMUP=1 timeout 120 python tools/quick_bench.py 522 1044 0 10 1 > gpurun_out/${1:-decompose3}_trace.log 2>&1
cat $out
