#!/bin/bash
# round-2 call 24 (re-run after the last kernel change): full GPU suite, final bench lines of every BASELINE config (with extras) and the launch list of the default bench command
set -u
O=gpurun_out/prof3; mkdir -p $O

python bench.py --steps 20 --warmup 3 2>$O/err_cfg2.log | tail -1 > $O/r2_bench_n1_resident.json; echo "cfg2 rc=$?"
python bench.py --config 3a --steps 80 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg3a.json
python bench.py --config 3b --steps 80 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg3b.json
python bench.py --config 4 --steps 3 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg4.json
python bench.py --config 5 --steps 300 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_cfg5.json
python bench.py --batch 512 --steps 40 --warmup 3 --no-extras 2>/dev/null | tail -1 > $O/r2_bench_cfg2_512_per_gpu.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/r2_bench_reference_arm.json
echo "bench lines done"
python bench.py --steps 2 --warmup 1 --no-extras > /dev/null 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_resident.csv python bench.py --steps 2 --warmup 1 --no-extras > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for c in 2 3a; do python tools/e2e_breakdown.py --config $c > $O/r2_e2e_breakdown_cfg$c.txt 2>&1; done
python - <<'PY'
import json
for n in ('r2_bench_n1_resident','r2_bench_cfg3a','r2_bench_cfg3b','r2_bench_cfg4','r2_bench_cfg5','r2_bench_cfg2_512_per_gpu','r2_bench_reference_arm'):
    j=json.loads(open('gpurun_out/prof3/%s.json'%n).read().strip().splitlines()[-1])
    r=j.get('roofline') or {}
    print(n,'value %.4g e2e %.4g host_loop %s frac %s clocks %s'%(j['value'],j['e2e']['value'],(j.get('host_loop') or {}).get('value'),r.get('frac'),(j.get('clocks') or {}).get('samples')))
PY
