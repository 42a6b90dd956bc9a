#!/bin/bash
# Collects the round-2 bench lines and ncu evidence kept under profiles/ (run on the GPU box through gpurun, 1 GPU).
# Everything lands in gpurun_out/prof as text (the .ncu-rep files stay on the box: gpurun merges back at most 64 MiB).
set -u
O=gpurun_out/prof; mkdir -p $O
T=/tmp/sbprof; mkdir -p $T
python bench.py --steps 20 --warmup 3 2>$O/err_cfg2.log | tail -1 > $O/r2_bench_n1_resident.json; echo "cfg2 rc=$?"
python bench.py --steps 8 --warmup 3 --design fused_tma --no-extras 2>/dev/null | tail -1 > $O/r2_bench_n1_fused_tma.json
python bench.py --steps 3 --warmup 3 --design split --no-extras 2>/dev/null | tail -1 > $O/r2_bench_n1_split.json
python bench.py --config 3a --steps 80 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg3a.json
python bench.py --config 3b --steps 80 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg3b.json
python bench.py --config 4 --steps 3 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg4.json
python bench.py --config 5 --steps 200 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg5.json
python bench.py --batch 512 --steps 40 --warmup 3 --no-extras 2>/dev/null | tail -1 > $O/r2_bench_cfg2_512_per_gpu.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/r2_bench_reference_arm.json
echo "bench lines done"
# launch list of the default bench command (plain run first)
python bench.py --steps 2 --warmup 1 --no-extras > /dev/null 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_resident.csv python bench.py --steps 2 --warmup 1 --no-extras > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# DRAM traffic of whole-solve launches: headline batch and a batch that does not fit the L2
for B in 4096 16384; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_solve_resident -c 2 --csv \
      --log-file $O/r2_dram_resident_B$B.csv python tools/sweep.py --config 2 --batch $B --reps 1 > $O/ncu_dram_$B.log 2>&1; echo "ncu dram $B rc=$?"
done
for c in 3a 4; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_solve_resident -c 2 --csv \
      --log-file $O/r2_dram_resident_cfg$c.csv python tools/sweep.py --config $c --reps 1 > $O/ncu_dram_$c.log 2>&1; echo "ncu dram cfg$c rc=$?"
done
# full captures (8 / 30 time steps keep the replays short); only the text pages travel
full() { # name cfg extra...
  local name=$1 cfg=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_resident -c 1 -o $T/$name -f \
      python tools/sweep.py --config $cfg --reps 1 "$@" > $O/ncu_full_$name.log 2>&1; echo "ncu full $name rc=$?"
  ncu -i $T/$name.ncu-rep --page raw --csv > $O/r2_ncu_full_${name}_raw.csv 2>/dev/null
  ncu -i $T/$name.ncu-rep --page source --csv > $O/r2_ncu_full_${name}_source.csv 2>/dev/null
}
full resident_cfg2 2 --time-steps 8
full resident_cfg3a 3a --time-steps 12
full resident_cfg4 4 --time-steps 30 --batch 4096
# long-mesh path: launch list with DRAM bytes, and the warm-stream timeline
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
    --log-file $O/r2_launches_longmesh.csv python bench.py --config 5 --steps 1 --warmup 1 --no-extras > $O/ncu_longmesh.log 2>&1; echo "ncu longmesh rc=$?"
python tools/trace_kernels.py --config 5 > $O/r2_trace_longmesh.txt 2>&1
du -sh $O
