#!/bin/bash
tag=${1:-r2n}
P=$PWD/image-lens-reproject_b200
for lib in liblrp_wta.so liblrp_wtb.so liblrp.so liblrp_wta.so; do
  LRP_LIB=$P/$lib timeout 300 python bench.py --steps 20 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --variant staged 2>gpurun_out/err_$tag.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib c2 fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'], 'clk', d['clocks']['sm_mhz'])"
done
