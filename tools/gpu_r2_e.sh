#!/bin/bash
# staged kernel: padded record rows A/B (LRP_REC_PAD), parity first
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "staged or c2_full or full_size or png_path or exr_path or tiled" 2>&1 | tail -6 > gpurun_out/pytest_$tag.log; tail -3 gpurun_out/pytest_$tag.log
for pad in 0; do
  echo "== staged LRP_REC_PAD=$pad" | tee -a gpurun_out/variants_$tag.jsonl
  LRP_REC_PAD=$pad timeout 300 python bench.py --steps 10 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --variant staged 2>gpurun_out/err_$tag.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 value', round(d['value'],2), 'us', round(d['roofline']['us_per_launch'],1), 'fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])" | tee -a gpurun_out/variants_$tag.jsonl
  LRP_REC_PAD=$pad timeout 600 python tools/bench_configs.py --configs c1t,c3,c4t,c5e,c5p --variants staged --coords table 2>>gpurun_out/err_$tag.txt | tee -a gpurun_out/variants_$tag.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(' ', d['config'], d['variant'], d['coords'], d['us_per_frame'])"
done
echo "== tiled v1" | tee -a gpurun_out/variants_$tag.jsonl
for ct in 1 0; do
LRP_TL_CTAS=$ct timeout 300 python bench.py --steps 10 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --variant tiled 2>>gpurun_out/err_$tag.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 tiled ct $ct value', round(d['value'],2), 'us', round(d['roofline']['us_per_launch'],1), 'fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])" | tee -a gpurun_out/variants_$tag.jsonl
done
