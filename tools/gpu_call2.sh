#!/bin/bash
# round-2 call 2: full GPU test suite with the fixed kernel, A/B of the second-level variants, ncu of a full launch
set -u
O=gpurun_out/c2; mkdir -p $O
L=skeelberzins.jl_b200
( time python -m pytest tests -m gpu -q --maxfail=6 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 $O/pytest_gpu.log
sw() { # lib-suffix tag cfg R T extra...
  local suf=$1 tag=$2 cfg=$3 R=$4 T=$5; shift 5
  [ -n "$suf" ] && export SB_LIB_PATH=$PWD/$L/libskeelberzins_b200$suf.so || unset SB_LIB_PATH
  if [ "$R" != "0" ]; then export SB_ROWS_PER_THREAD=$R SB_THREADS_PER_SYSTEM=$T; else unset SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM; fi
  timeout 300 python tools/sweep.py --config $cfg --tag $tag "$@" 2>>$O/sweep.err | tail -1 | tee -a $O/sweep.jsonl
}
sw "" new 2 0 0 --check
sw _ser ser 2 0 0 --check
sw _rcp rcp 2 0 0 --check
sw "" new1s 2 0 0 --streams 1
sw _ser ser1s 2 0 0 --streams 1
sw "" new 3a 0 0 --check
sw _ser ser 3a 0 0 --check
sw _rcp rcp 3a 0 0
sw "" new 3a 16 32
sw "" new 3b 0 0 --check
sw "" new 4 0 0 --check
sw "" new16k 2 0 0 --batch 16384 --time-steps 12 --streams 1
sw _ser ser16k 2 0 0 --batch 16384 --time-steps 12 --streams 1
sw _r1 r116k 2 0 0 --batch 16384 --time-steps 12 --streams 1
unset SB_LIB_PATH SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_newton_fused_p -s 2 -c 2 -o $O/fused_p_new -f \
   python tools/sweep.py --config 2 --streams 1 --time-steps 3 --reps 1 > $O/ncu_full.log 2>&1; echo "ncu rc=$?"
export SB_LIB_PATH=$PWD/$L/libskeelberzins_b200_ser.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_newton_fused_p -s 2 -c 2 -o $O/fused_p_ser -f \
   python tools/sweep.py --config 2 --streams 1 --time-steps 3 --reps 1 > $O/ncu_full_ser.log 2>&1; echo "ncu ser rc=$?"
ls -la $O
