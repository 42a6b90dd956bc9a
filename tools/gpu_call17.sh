#!/bin/bash
# round-2 call 17: mesh weights on the device for m == 0; full GPU suite; set-up time of cfg5
set -u
O=gpurun_out/c19; mkdir -p $O
( time python -m pytest tests -m gpu -q --maxfail=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
python tools/e2e_breakdown.py > $O/e2e_breakdown_cfg5.txt 2>&1; cat $O/e2e_breakdown_cfg5.txt
python bench.py --config 5 --steps 300 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg5.json
python - <<'PY'
import json
j=json.loads(open('gpurun_out/c19/r2_bench_cfg5.json').read().strip().splitlines()[-1])
print('cfg5 value',j['value'],'e2e',j['e2e']['value'],'ms_e2e',j['e2e']['ms_per_step'],'us/iter',j['roofline']['avg_launch_us'],'frac',j['roofline']['frac'])
PY
