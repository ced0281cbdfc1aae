#!/bin/bash
# round 2, step w: warps per CTA of the 3-channel staged bicubic kernel (20 / 24 / 28) after the record-layout change
cd /root/repo
P=/root/repo/image-lens-reproject_b200
for lib in liblrp.so liblrp_w24.so liblrp_w28.so; do
  LRP_LIB=$P/$lib timeout 300 python bench.py --steps 20 --quick --no-cpu-baseline --no-sched --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib c2 fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'], 'value us', round(d['roofline']['us_per_launch'],2))"
  for cm in table fly; do
  LRP_LIB=$P/$lib timeout 600 python tools/bench_configs.py --configs c1t,c5e,c5p --variants staged --coords $cm 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(' ', d['config'], d['coords'], d['us_per_frame'])"
  done
done
