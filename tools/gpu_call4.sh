#!/bin/bash
# round-2 call 4: full tests after the alignment fix, variants of the whole-solve kernel, first bench lines
set -u
O=gpurun_out/c4; mkdir -p $O
L=skeelberzins.jl_b200
( time python -m pytest tests -m gpu -q --maxfail=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 $O/pytest_gpu.log
sw() { # lib-suffix tag cfg R T extra...
  local suf=$1 tag=$2 cfg=$3 R=$4 T=$5; shift 5
  [ -n "$suf" ] && export SB_LIB_PATH=$PWD/$L/libskeelberzins_b200$suf.so || unset SB_LIB_PATH
  if [ "$R" != "0" ]; then export SB_ROWS_PER_THREAD=$R SB_THREADS_PER_SYSTEM=$T; else unset SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM; fi
  timeout 300 python tools/sweep.py --config $cfg --tag $tag "$@" 2>>$O/sweep.err | tail -1 | tee -a $O/sweep.jsonl
}
sw "" res 2 0 0
sw _ser res_ser 2 0 0 --check
sw _nomon res_nomon 2 0 0
sw "" res 3a 0 0
sw _ser res_ser 3a 0 0
sw _nomon res_nomon 3a 0 0
sw "" res 4 0 0
sw "" res 4 4 64 --check
sw "" tma 4 0 0 --design fused_tma
sw "" res 5 0 0
unset SB_LIB_PATH SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM
python bench.py --steps 20 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"; tail -c 3000 $O/bench_cfg2.json
python bench.py --config 3a --steps 60 --warmup 3 > $O/bench_cfg3a.json 2> $O/bench_cfg3a.err; echo "bench3a rc=$?"
python bench.py --config 4 --steps 3 --warmup 3 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench4 rc=$?"
ls -la $O
