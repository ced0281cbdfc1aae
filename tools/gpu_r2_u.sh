#!/bin/bash
P=$PWD/image-lens-reproject_b200
for lib in liblrp.so liblrp_u6.so liblrp_u8.so; do
  LRP_LIB=$P/$lib timeout 300 python bench.py --steps 20 --quick --no-cpu-baseline --no-sched --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib c2 fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])"
  LRP_LIB=$P/$lib timeout 600 python tools/bench_configs.py --configs c1t,c3,c4t,c5e --variants staged --coords table 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(' ', d['config'], d['coords'], d['us_per_frame'])"
done
