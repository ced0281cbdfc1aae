#!/bin/bash
# round 2, GPU cycle c: tiled kernel parity after the threshold fix, tile-shape A/B libs, nn table kernel timing
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "tiled or c1t or c3_full or pole or nearest or c2_full" 2>&1 | tail -15 > gpurun_out/pytest_tiled_$tag.log; tail -4 gpurun_out/pytest_tiled_$tag.log
rm -f gpurun_out/variants_$tag.jsonl
P=$PWD/image-lens-reproject_b200
for lib in liblrp.so liblrp_w4r4.so; do
  for ct in 1 0; do
    echo "== $lib ctas-choice $ct" | tee -a gpurun_out/variants_$tag.jsonl
    LRP_LIB=$P/$lib LRP_TL_CTAS=$ct timeout 300 python bench.py --steps 10 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --variant tiled 2>gpurun_out/err_$tag.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 value', round(d['value'],2), 'us', round(d['roofline']['us_per_launch'],1), 'fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])" | tee -a gpurun_out/variants_$tag.jsonl
    LRP_LIB=$P/$lib LRP_TL_CTAS=$ct timeout 600 python tools/bench_configs.py --configs c1t,c3,c4t,c5e --variants tiled --coords table 2>>gpurun_out/err_$tag.txt | tee -a gpurun_out/variants_$tag.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(' ', d['config'], d['variant'], d['coords'], d['us_per_frame'])"
  done
done
timeout 300 python bench.py --steps 10 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --interp nn 2>>gpurun_out/err_$tag.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nn value', round(d['value'],2), 'us', round(d['roofline']['us_per_launch'],1), 'frac', round(d['roofline']['frac'],4), 'legs', d['coords_legs'])" | tee -a gpurun_out/variants_$tag.jsonl
