#!/bin/bash
# round-2 call 26: PIPE template variant (the headline kernel keeps its old code); ring fix; full suite; values
set -u
O=gpurun_out/c27; mkdir -p $O
( time timeout -s KILL 300 python -m pytest tests/test_gpu_resident.py -m gpu -q --maxfail=4 ) > $O/pytest_resident.log 2>&1; echo "pytest resident rc=$?"; tail -3 $O/pytest_resident.log
( time timeout -s KILL 400 python -m pytest tests -m gpu -q --maxfail=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
for c in 2 3a 3b 4; do
  st=20; [ "$c" = "3a" ] && st=80; [ "$c" = "3b" ] && st=80; [ "$c" = "4" ] && st=3
  timeout -s KILL 300 python bench.py --config $c --steps $st --warmup 3 --no-extras 2>/dev/null | tail -1 > $O/bench_cfg$c.json
  C=$c python - <<'PY'
import json,os
c=os.environ['C']
j=json.loads(open('gpurun_out/c27/bench_cfg%s.json'%c).read().strip().splitlines()[-1])
print('cfg',c,'value %.4g e2e %.4g ms_e2e %.3f ms_step %.3f'%(j['value'],j['e2e']['value'],j['e2e']['ms_per_step'],j['ms_per_step']))
PY
done
