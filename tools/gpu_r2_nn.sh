#!/bin/bash
for per in 4 8 16; do
LRP_NN_PER=$per timeout 300 python bench.py --steps 30 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --interp nn 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nn per $per value', round(d['value'],2), 'frac', round(d['roofline']['frac'],4), 'table', d['coords_legs']['table'])"
done
