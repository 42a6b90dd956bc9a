#!/bin/bash
# round-2 call 28 (8 GPUs): weak and strong line of the headline workload with the final library (pipelined e2e leg on 8 ranks)
set -u
O=gpurun_out/scale2; mkdir -p $O
run() { # N scaling steps
  local N=$1 sc=$2 steps=$3
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps $steps --warmup 3 --scaling $sc --no-extras 2>$O/err_${sc}_n$N.log | tail -1 > $O/r2_bench_${sc}_n$N.json
  echo "$sc N=$N rc=$? $(python -c "import json,sys; d=json.load(open('$O/r2_bench_${sc}_n$N.json')); print('%.4g units/s e2e %.4g  %.3f ms/step  per-rank %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['per_rank_ms_per_step']['all']))" 2>&1 | tail -1)"
}
run 8 weak 20
run 8 strong 160
run 4 weak 20
