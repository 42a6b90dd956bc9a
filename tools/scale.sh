#!/bin/bash
# Weak and strong scaling lines of the headline workload on one box (run through `gpurun --gpus 8`).
#   weak:   4096 instances per GPU (what the driver's 1,2,4,8 sweep runs)
#   strong: 4096 instances in total, 4096/N per GPU (512 per GPU at N = 8)
set -u
O=gpurun_out/scale; mkdir -p $O
run() { # N scaling steps
  local N=$1 sc=$2 steps=$3
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps $steps --warmup 3 --scaling $sc --no-extras 2>$O/err_${sc}_n$N.log | tail -1 > $O/r2_bench_${sc}_n$N.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
        bench.py --gpus $N --steps $steps --warmup 3 --scaling $sc --no-extras 2>$O/err_${sc}_n$N.log | tail -1 > $O/r2_bench_${sc}_n$N.json
  fi
  echo "$sc N=$N rc=$? $(python -c "import json,sys; d=json.load(open('$O/r2_bench_${sc}_n$N.json')); print('%.4g units/s  %.3f ms/step  per-rank %s' % (d['value'], d['ms_per_step'], d['config']['per_rank_ms_per_step']['all']))" 2>&1 | tail -1)"
}
nvidia-smi topo -m > $O/topology.txt 2>&1
run 1 weak 20
for N in 2 4 8; do run $N weak 20; done
for N in 2 4 8; do run $N strong $((20 * N)); done
ls -la $O
