#!/bin/bash
# Thread-decomposition sweep for one BASELINE config: tools/sweep_decomp.sh <cfg> "<R> <TT>" ["<R> <TT>" ...]
cfg=$1; shift
for rt in "$@"; do
set -- $rt
SB_ROWS_PER_THREAD=$1 SB_THREADS_PER_SYSTEM=$2 python bench.py --config $cfg --steps 2 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg$cfg', d['config'].get('rows_per_thread'), d['config'].get('threads_per_system'), '%.3e'%d['value'], '%.2f ms'%d['ms_per_step'], '%.1f us/launch'%d['roofline']['avg_launch_us'], 'frac %.3f'%d['roofline']['frac'])"
done
