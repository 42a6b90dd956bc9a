#!/bin/bash
# round-2 call 19: e2e breakdown of the transient configs
set -u
O=gpurun_out/c21; mkdir -p $O
for c in 3a 2 4; do echo "== cfg$c"; timeout 300 python tools/e2e_breakdown.py --config $c 2>&1 | tee $O/e2e_breakdown_cfg$c.txt; done
