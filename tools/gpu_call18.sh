#!/bin/bash
# round-2 call 18: systems (npde >= 2) on long meshes
set -u
O=gpurun_out/c20; mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin.py -m gpu -q --maxfail=12 -k "longmesh or long_mesh or cfg5 or wide_systems" ) > $O/pytest_long.log 2>&1; echo "pytest long rc=$?"
tail -30 $O/pytest_long.log
