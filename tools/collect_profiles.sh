#!/bin/bash
# Collects the bench lines and ncu evidence kept under profiles/ (run on the GPU box through gpurun).
set -u
mkdir -p gpurun_out/prof
python bench.py --steps 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/prof/r1_bench_n1_fused_tma.json; echo "default rc=$?"
python bench.py --steps 3 --warmup 3 --design split 2>/dev/null | tail -1 > gpurun_out/prof/r1_bench_n1_split.json
python bench.py --steps 3 --warmup 3 --design fused 2>/dev/null | tail -1 > gpurun_out/prof/r1_bench_n1_fused.json
for c in 3a 3b 4 5; do python bench.py --config $c --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/prof/r1_bench_cfg$c.json; done
python bench.py --steps 1 --warmup 1 --time-steps 6 > /dev/null 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/launches_fused_tma.csv python bench.py --steps 1 --warmup 1 --time-steps 6 > gpurun_out/prof/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# full-batch launches (one problem, as in the roofline pass of bench.py): --streams 1
ncu --set full --clock-control none --import-source on -k regex:k_newton_fused_p -s 6 -c 3 -o gpurun_out/prof/fused_p_full -f python bench.py --steps 1 --warmup 1 --time-steps 6 --streams 1 > gpurun_out/prof/ncu_full.log 2>&1; echo "ncu full rc=$?"
