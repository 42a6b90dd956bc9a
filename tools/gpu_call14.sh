#!/bin/bash
# round-2 call 14: ncu full captures: whole-solve kernel on cfg3a (two systems per CTA), long-mesh sweeps (final structure)
set -u
O=gpurun_out/c17; mkdir -p $O
T=/tmp/sbprof; mkdir -p $T
cap() { # name kernel-regex skip cfg extra...
  local name=$1 rx=$2 skip=$3 cfg=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $T/$name -f \
      python tools/sweep.py --config $cfg --reps 1 "$@" > $O/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i $T/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
  ncu -i $T/$name.ncu-rep --page source --csv > $O/${name}_source.csv 2>/dev/null
}
cap resident_cfg3a k_solve_resident 1 3a --time-steps 12
cap long_fwd k6_chunk_asm 8 5
cap long_bwd k6_back_asm 8 5
ls -la $O
