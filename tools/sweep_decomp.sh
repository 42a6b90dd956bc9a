for cfg in 3a 3b; do
for rt in "8 64" "16 32" "4 128" "16 64"; do
set -- $rt
SB_ROWS_PER_THREAD=$1 SB_THREADS_PER_SYSTEM=$2 python bench.py --config $cfg --steps 2 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', d['config'].get('rows_per_thread'), d['config'].get('threads_per_system'), '%.3e'%d['value'], d['ms_per_step'], d['roofline']['avg_launch_us'], d['roofline']['frac'])"
done; done
