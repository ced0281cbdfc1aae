#!/bin/bash
# round 2, GPU cycle b: the tiled kernel — parity, then timing against staged, then ncu of the nn table kernel
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "tiled or c1t or c3_full or pole" 2>&1 | tail -30 > gpurun_out/pytest_tiled_$tag.log; tail -5 gpurun_out/pytest_tiled_$tag.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu_$tag.log; tail -5 gpurun_out/pytest_gpu_$tag.log
rm -f gpurun_out/variants_$tag.jsonl
for v in staged tiled; do
  for ct in 3 2; do
    [ $v = staged ] && [ $ct = 2 ] && continue
    LRP_TL_CTAS=$ct timeout 300 python bench.py --steps 10 --quick --no-cpu-baseline --no-sched --e2e-steps 2 --variant $v 2>gpurun_out/err_$tag.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v ctas $ct value', round(d['value'],2), 'us', round(d['roofline']['us_per_launch'],1), 'fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])" | tee -a gpurun_out/variants_$tag.jsonl
    LRP_TL_CTAS=$ct timeout 600 python tools/bench_configs.py --configs c1t,c3,c4t,c5e,c5p --variants $v --coords fly,table 2>>gpurun_out/err_$tag.txt | tee -a gpurun_out/variants_$tag.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(' ', d['config'], d['variant'], d['coords'], d['us_per_frame'])"
  done
done
