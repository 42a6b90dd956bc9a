#!/bin/bash
# round-2 call 3: the whole-solve kernel -- tests, timing against fused_tma, variants, ncu
set -u
O=gpurun_out/c3; mkdir -p $O
L=skeelberzins.jl_b200
( time python -m pytest tests -m gpu -q --maxfail=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 $O/pytest_gpu.log
sw() { # lib-suffix tag cfg R T extra...
  local suf=$1 tag=$2 cfg=$3 R=$4 T=$5; shift 5
  [ -n "$suf" ] && export SB_LIB_PATH=$PWD/$L/libskeelberzins_b200$suf.so || unset SB_LIB_PATH
  if [ "$R" != "0" ]; then export SB_ROWS_PER_THREAD=$R SB_THREADS_PER_SYSTEM=$T; else unset SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM; fi
  timeout 300 python tools/sweep.py --config $cfg --tag $tag "$@" 2>>$O/sweep.err | tail -1 | tee -a $O/sweep.jsonl
}
sw "" res 2 0 0 --check
sw "" tma 2 0 0 --design fused_tma
sw _ser res_ser 2 0 0 --check
sw _nomon res_nomon 2 0 0
sw "" res 3a 0 0 --check
sw "" tma 3a 0 0 --design fused_tma
sw _ser res_ser 3a 0 0
sw "" res 3a 16 32
sw "" res 3b 0 0 --check
sw "" res 4 0 0 --check
sw "" tma 4 0 0 --design fused_tma
sw _nomon res_nomon 4 0 0
sw "" res512 2 0 0 --batch 512
sw "" tma512 2 0 0 --batch 512 --design fused_tma
sw "" res16k 2 0 0 --batch 16384 --time-steps 20
sw "" res 2 16 64
sw "" res 2 4 256
SB_PERSISTENT_CTAS_PER_SM=3 sw "" res_3cta 2 0 0
SB_PERSISTENT_CTAS_PER_SM=2 sw "" res_2cta 2 0 0
unset SB_LIB_PATH SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_resident -c 1 -o $O/resident -f \
   python tools/sweep.py --config 2 --time-steps 8 --reps 1 > $O/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la $O
