#!/bin/bash
# round-2 call 7: npde = 2 (cfg4) under register caps of 255 / 168 / 128 per thread x decompositions (8,32) / (4,64)
set -u
O=gpurun_out/c10; mkdir -p $O
sw() { # tag lib cfg R T extra...
  local tag=$1 lib=$2 cfg=$3 R=$4 T=$5; shift 5
  export SB_LIB_PATH=$PWD/skeelberzins.jl_b200/libskeelberzins_b200$lib.so
  if [ "$R" != "0" ]; then export SB_ROWS_PER_THREAD=$R SB_THREADS_PER_SYSTEM=$T; else unset SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM; fi
  timeout 300 python tools/sweep.py --config $cfg --tag $tag "$@" 2>>$O/sweep.err | tail -1 | tee -a $O/sweep.jsonl
}
sw p8 "" 4 8 32 --check
sw p8 "" 4 4 64
sw p12 _p12 4 8 32 --check
sw p12 _p12 4 4 64 --check
sw p16 _p16 4 8 32
sw p16 _p16 4 4 64 --check
tail -5 $O/sweep.err
