#!/bin/bash
# round-2 call 21 (2 GPUs): weak and strong lines of the headline workload with the final library; the sharded GPU tests
set -u
O=gpurun_out/c23; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 20 --warmup 3 --no-extras 2>$O/err_weak_n2.log | tail -1 > $O/r2_bench_weak_n2.json; echo "weak rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 \
    bench.py --gpus 2 --steps 40 --warmup 3 --scaling strong --no-extras 2>$O/err_strong_n2.log | tail -1 > $O/r2_bench_strong_n2.json; echo "strong rc=$?"
python - <<'PY'
import json
for n in ('weak','strong'):
    d=json.load(open('gpurun_out/c23/r2_bench_%s_n2.json'%n))
    print(n,'value %.4g e2e %.4g ms/step %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d['config'].get('per_rank_ms_per_step'))
PY
( timeout 600 python -m pytest tests/test_gpu_resident.py -m gpu -q -k "devices or final_state" ) > $O/pytest_devices.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_devices.log
