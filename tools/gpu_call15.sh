#!/bin/bash
# round-2 call 15: long-mesh level-0 kernels compiled for 16 (default) / 20 / 24 warps per SM
set -u
O=gpurun_out/c18; mkdir -p $O
for v in "" _w20 _w24; do
  export SB_LIB_PATH=$PWD/skeelberzins.jl_b200/libskeelberzins_b200$v.so
  ( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=8 -k "longmesh or cfg5" ) > $O/pytest_long$v.log 2>&1; echo "pytest $v rc=$?"
  timeout 300 python bench.py --config 5 --steps 100 --warmup 3 --no-extras > $O/bench_cfg5$v.json 2> $O/bench_cfg5$v.err; echo "bench5 $v rc=$?"
  V=$v python - <<'PY'
import json,os
v=os.environ['V']
j=json.loads(open('gpurun_out/c18/bench_cfg5%s.json'%v).read().strip().splitlines()[-1])
print('variant',v or 'w16','value',j['value'],'e2e',j['e2e']['value'],'us/iter',j['roofline']['avg_launch_us'])
PY
  timeout 300 python tools/trace_kernels.py --config 5 > $O/trace_longmesh$v.txt 2>&1; grep k6_ $O/trace_longmesh$v.txt | tail -14 | head -4
done
