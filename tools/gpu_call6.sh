#!/bin/bash
# round-2 call 5: serial second level as default + monitor, full tests, sweeps, ncu of the whole-solve kernel, bench lines
set -u
O=gpurun_out/c6; mkdir -p $O
( time python -m pytest tests -m gpu -q --maxfail=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 $O/pytest_gpu.log
sw() { # tag cfg R T extra...
  local tag=$1 cfg=$2 R=$3 T=$4; shift 4
  if [ "$R" != "0" ]; then export SB_ROWS_PER_THREAD=$R SB_THREADS_PER_SYSTEM=$T; else unset SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM; fi
  timeout 300 python tools/sweep.py --config $cfg --tag $tag "$@" 2>>$O/sweep.err | tail -1 | tee -a $O/sweep.jsonl
}
sw res 2 0 0 --check
sw res 3a 0 0 --check
sw res 3b 0 0 --check
sw res 4 0 0 --check
sw res 2 16 64
sw res 3a 16 32
sw res 3a 4 128
sw res512 2 0 0 --batch 512
sw res1024 2 0 0 --batch 1024
sw res2048 2 0 0 --batch 2048
sw res16k 2 0 0 --batch 16384 --time-steps 20
sw tma 2 0 0 --design fused_tma
unset SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM
python bench.py --steps 20 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"
ls -la $O
