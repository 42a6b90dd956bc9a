#!/bin/bash
# Round-2 evidence for what changed after tools/collect_profiles.sh ran (long-mesh path, two systems per CTA at TT = 64):
# bench lines of every BASELINE config, launch list + DRAM bytes + warm-stream trace of cfg5, ncu full captures of the
# long-mesh sweeps and of the whole-solve kernel on cfg3a.  1 GPU; everything lands in gpurun_out/prof2 as text.
set -u
O=gpurun_out/prof2; mkdir -p $O
T=/tmp/sbprof; mkdir -p $T
( time python -m pytest tests -m gpu -q --maxfail=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 3 2>$O/err_cfg2.log | tail -1 > $O/r2_bench_n1_resident.json; echo "cfg2 rc=$?"
python bench.py --config 3a --steps 80 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg3a.json
python bench.py --config 3b --steps 80 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg3b.json
python bench.py --config 4 --steps 3 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg4.json
python bench.py --config 5 --steps 300 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg5.json
echo "bench lines done"
python tools/e2e_breakdown.py > $O/r2_e2e_breakdown_cfg5.txt 2>&1
# long-mesh path: launch list with DRAM bytes (plain run first), warm-stream timeline
python bench.py --config 5 --steps 1 --warmup 1 --no-extras > /dev/null 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
    --log-file $O/r2_launches_longmesh.csv python bench.py --config 5 --steps 1 --warmup 1 --no-extras > $O/ncu_longmesh.log 2>&1; echo "ncu longmesh rc=$?"
python tools/trace_kernels.py --config 5 > $O/r2_trace_longmesh.txt 2>&1
cap() { # name kernel-regex skip cfg extra...
  local name=$1 rx=$2 skip=$3 cfg=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $T/$name -f \
      python tools/sweep.py --config $cfg --reps 1 "$@" > $O/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i $T/$name.ncu-rep --page raw --csv > $O/r2_ncu_full_${name}_raw.csv 2>/dev/null
}
cap resident_cfg3a k_solve_resident 1 3a --time-steps 12
cap longmesh_fwd k6_chunk_asm 8 5
cap longmesh_bwd k6_back_asm 8 5
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_solve_resident -c 2 --csv \
    --log-file $O/r2_dram_resident_cfg3a.csv python tools/sweep.py --config 3a --reps 1 > $O/ncu_dram_3a.log 2>&1; echo "ncu dram cfg3a rc=$?"
du -sh $O
