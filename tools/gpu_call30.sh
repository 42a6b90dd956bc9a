#!/bin/bash
# round-2 call 30: 2 MB chunks, cached time grids: resident tests, e2e of cfg2 / cfg3a / cfg4
set -u
O=gpurun_out/c30; mkdir -p $O
( time timeout -s KILL 300 python -m pytest tests/test_gpu_resident.py -m gpu -q --maxfail=4 ) > $O/pytest_resident.log 2>&1; echo "pytest resident rc=$?"; tail -3 $O/pytest_resident.log
for c in 2 3a 4; do
  st=20; [ "$c" = "3a" ] && st=80; [ "$c" = "4" ] && st=3
  timeout -s KILL 300 python bench.py --config $c --steps $st --warmup 3 --no-extras 2>/dev/null | tail -1 > $O/bench_cfg$c.json
  C=$c python - <<'PY'
import json,os
c=os.environ['C']
j=json.loads(open('gpurun_out/c30/bench_cfg%s.json'%c).read().strip().splitlines()[-1])
print('cfg',c,'value %.4g e2e %.4g ms_e2e %.3f ms_step %.3f'%(j['value'],j['e2e']['value'],j['e2e']['ms_per_step'],j['ms_per_step']))
PY
done
