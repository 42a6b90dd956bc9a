#!/bin/bash
# round-2 call 10: ncu full capture of the two long-mesh sweeps (T0 = 128 build)
set -u
O=gpurun_out/c13; mkdir -p $O
T=/tmp/sbprof; mkdir -p $T
export SB_LIB_PATH=$PWD/skeelberzins.jl_b200/libskeelberzins_b200_t128.so
python tools/sweep.py --config 5 --reps 1 > $O/plain.log 2>&1; echo "plain rc=$?"; tail -1 $O/plain.log
for kname in k6_chunk_asm k6_back_asm; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kname -s 8 -c 1 -o $T/$kname -f \
      python tools/sweep.py --config 5 --reps 1 > $O/ncu_$kname.log 2>&1; echo "ncu $kname rc=$?"
  ncu -i $T/$kname.ncu-rep --page raw --csv > $O/${kname}_raw.csv 2>/dev/null
  ncu -i $T/$kname.ncu-rep --page source --csv > $O/${kname}_source.csv 2>/dev/null
done
ls -la $O
