#!/bin/bash
# round-2 call 13: full GPU suite; two 64-thread systems per CTA in the whole-solve kernel (cfg3); long-mesh prologue reordered
set -u
O=gpurun_out/c16; mkdir -p $O
( time python -m pytest tests -m gpu -q --maxfail=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 $O/pytest_gpu.log
sw() { # tag cfg R T extra...
  local tag=$1 cfg=$2 R=$3 T=$4; shift 4
  if [ "$R" != "0" ]; then export SB_ROWS_PER_THREAD=$R SB_THREADS_PER_SYSTEM=$T; else unset SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM; fi
  timeout 300 python tools/sweep.py --config $cfg --tag $tag "$@" 2>>$O/sweep.err | tail -1 | tee -a $O/sweep.jsonl
}
sw res 2 0 0
sw res 3a 0 0 --check
sw res 3b 0 0 --check
sw res 4 0 0
sw res64 2 8 64 --check
unset SB_ROWS_PER_THREAD SB_THREADS_PER_SYSTEM
timeout 300 python bench.py --config 5 --steps 100 --warmup 3 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench5 rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c16/bench_cfg5.json').read().strip().splitlines()[-1])
print('cfg5 value',j['value'],'e2e',j['e2e']['value'],'ms_e2e',j['e2e']['ms_per_step'],'us/iter',j['roofline']['avg_launch_us'])
PY
timeout 300 python tools/trace_kernels.py --config 5 > $O/trace_longmesh.txt 2>&1; grep k6_ $O/trace_longmesh.txt | tail -14 | head -8
