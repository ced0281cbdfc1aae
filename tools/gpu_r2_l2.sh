#!/bin/bash
for mb in 0 72 100; do
for it in nn bc bl; do
LRP_L2_PERSIST_MB=$mb timeout 300 python bench.py --steps 30 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --interp $it 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('persist $mb MB $it value', round(d['value'],2), 'us', round(d['roofline']['us_per_launch'],2), 'frac', round(d['roofline']['frac'],4), 'table', d['coords_legs']['table']['us_per_launch'], 'fly', d['coords_legs']['fly']['us_per_launch'])"
done; done
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:nn_table -s 60 -c 2 python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --interp nn 2>&1 | grep -E "dram__|lts__|gpu__time" | head -8
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:reproject_staged -s 60 -c 2 python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline --no-sched --e2e-steps 1 2>&1 | grep -E "dram__|lts__|gpu__time" | head -8
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "coords or nearest or remap" 2>&1 | tail -2
