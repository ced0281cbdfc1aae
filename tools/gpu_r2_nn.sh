#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "nearest or coords_modes or c2_full_size_nearest" 2>&1 | tail -2
for c in 1 0; do
LRP_NN_COMPACT=$c timeout 300 python bench.py --steps 30 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --interp nn 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nn compact $c value', round(d['value'],2), 'frac', round(d['roofline']['frac'],4), 'table', d['coords_legs']['table'])"
done
