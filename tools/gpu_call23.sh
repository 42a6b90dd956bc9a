#!/bin/bash
# round-2 call 23: sb_solve_from_host with every copy enqueued before the launch; devices test; full suite under a hard timeout
set -u
O=gpurun_out/c25; mkdir -p $O
( time timeout -s KILL 200 python -m pytest tests/test_gpu_resident.py -m gpu -q --maxfail=4 -k "final_state or devices" ) > $O/pytest_final.log 2>&1; echo "pytest final rc=$?"; tail -25 $O/pytest_final.log
( time timeout -s KILL 400 python -m pytest tests -m gpu -q --maxfail=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
