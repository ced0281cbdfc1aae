#!/bin/bash
tag=${1:-r2o}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --quick --no-cpu-baseline --no-sched --e2e-steps 1 2>gpurun_out/err_$tag.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default c2 value', round(d['value'],2), 'fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_$tag.log; tail -5 gpurun_out/pytest_gpu_$tag.log
