#!/bin/bash
# round-2 call 1: parity of the new solve + variant sweep + one full ncu capture
set -u
O=gpurun_out/c1; mkdir -p $O
L=skeelberzins.jl_b200
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest_gpu.log
sw() { # lib-suffix tag cfg R T extra...
  local suf=$1 tag=$2 cfg=$3 R=$4 T=$5; shift 5
  local env=""
  [ -n "$suf" ] && export SB_LIB_PATH=$PWD/$L/libskeelberzins_b200$suf.so || unset SB_LIB_PATH
  if [ "$R" != "0" ]; then export SB_ROWS_PER_THREAD=$R SB_THREADS_PER_SYSTEM=$T; else unset SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM; fi
  timeout 300 python tools/sweep.py --config $cfg --tag $tag "$@" 2>>$O/sweep.err | tail -1 | tee -a $O/sweep.jsonl
}
sw _r1 r1 2 0 0
sw _r1 r1 3a 0 0
sw _r1 r1 4 0 0
sw "" new 2 0 0 --check
sw "" new 2 16 64 --check
sw "" new 2 4 256 --check
sw "" new 2 8 256
sw "" new 3a 0 0 --check
sw "" new 3a 16 32 --check
sw "" new 3a 4 128 --check
sw "" new 3b 0 0 --check
sw "" new 4 0 0 --check
sw "" new 4 4 64 --check
sw "" new1s 2 0 0 --streams 1
sw _s2 s2 2 0 0 --check
sw _s2 s2 3a 0 0
sw _w20 w20 2 0 0 --check
sw _w20 w20 3a 0 0
sw _w20 w20 2 16 64
unset SB_LIB_PATH SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM
# full ncu capture of the new cfg2 kernel (whole-batch launches, one stream)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_newton_fused_p -s 6 -c 2 -o $O/fused_p_new -f \
   python tools/sweep.py --config 2 --streams 1 --time-steps 4 --reps 1 > $O/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la $O
