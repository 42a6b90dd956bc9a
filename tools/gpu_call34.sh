#!/bin/bash
# round-2 call 34: ncu full captures of the final whole-solve kernel on cfg2 / cfg3a / cfg4 (first time steps only: the replays stay short)
set -u
O=gpurun_out/prof4; mkdir -p $O
T=/tmp/sbprof; mkdir -p $T
cap() { # name kernel-regex skip cfg extra...
  local name=$1 rx=$2 skip=$3 cfg=$4; shift 4
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $T/$name -f \
      python tools/sweep.py --config $cfg --reps 1 "$@" > $O/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i $T/$name.ncu-rep --page raw --csv > $O/r2_ncu_full_${name}_raw.csv 2>/dev/null
}
cap resident_cfg2 k_solve_resident 1 2 --time-steps 8
cap resident_cfg3a k_solve_resident 1 3a --time-steps 12
cap resident_cfg4 k_solve_resident 1 4 --time-steps 30 --batch 4096
ls -la $O
