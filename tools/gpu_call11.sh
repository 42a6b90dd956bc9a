#!/bin/bash
# round-2 call 11: persistent double-buffered level-0 kernels; tiles of 2048 (default) and 1024 nodes; e2e breakdown
set -u
O=gpurun_out/c14; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin.py -m gpu -q --maxfail=8 -k "longmesh or long_mesh or cfg5" ) > $O/pytest_long.log 2>&1; echo "pytest long rc=$?"
tail -5 $O/pytest_long.log
timeout 300 python bench.py --config 5 --steps 100 --warmup 3 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench5 rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c14/bench_cfg5.json').read().strip().splitlines()[-1])
print('T0=256 value',j['value'],'e2e',j['e2e']['value'],'ms_e2e',j['e2e']['ms_per_step'],'us/iter',j['roofline']['avg_launch_us'])
PY
timeout 300 python tools/trace_kernels.py --config 5 > $O/trace_longmesh.txt 2>&1; tail -22 $O/trace_longmesh.txt | head -12
export SB_LIB_PATH=$PWD/skeelberzins.jl_b200/libskeelberzins_b200_t128.so
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=8 -k "longmesh or cfg5" ) > $O/pytest_long_t128.log 2>&1; echo "pytest t128 rc=$?"
tail -3 $O/pytest_long_t128.log
timeout 300 python bench.py --config 5 --steps 100 --warmup 3 --no-extras > $O/bench_cfg5_t128.json 2> $O/bench_cfg5_t128.err; echo "bench5 t128 rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c14/bench_cfg5_t128.json').read().strip().splitlines()[-1])
print('T0=128 value',j['value'],'e2e',j['e2e']['value'],'us/iter',j['roofline']['avg_launch_us'])
PY
timeout 300 python tools/trace_kernels.py --config 5 > $O/trace_longmesh_t128.txt 2>&1; tail -22 $O/trace_longmesh_t128.txt | head -12
unset SB_LIB_PATH
timeout 300 python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1; cat $O/e2e_breakdown.txt
