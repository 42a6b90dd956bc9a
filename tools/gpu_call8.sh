#!/bin/bash
# round-2 call 8: long-mesh path with one summary per tile (in-CTA open reduction) and weights from x
set -u
O=gpurun_out/c11; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin.py -m gpu -q --maxfail=8 -k "longmesh or long_mesh or cfg5" ) > $O/pytest_long.log 2>&1; echo "pytest long rc=$?"
tail -15 $O/pytest_long.log
timeout 300 python bench.py --config 5 --steps 100 --warmup 3 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench5 rc=$?"
tail -c 1500 $O/bench_cfg5.json
timeout 300 python tools/trace_kernels.py --config 5 > $O/trace_longmesh.txt 2>&1; tail -30 $O/trace_longmesh.txt
