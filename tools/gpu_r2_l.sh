#!/bin/bash
# cp.async staging A/B (LRP_STAGE_ASYNC): parity with the switch on, then c2 / c1t / c3 / c4t / c5e / c5p timings both ways
tag=${1:-r2l}
mkdir -p gpurun_out
LRP_STAGE_ASYNC=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "staged or c2_full or full_size or png_path or exr_path or pole or c1t or c3_full or supersampling_staged" 2>&1 | tail -4 > gpurun_out/pytest_$tag.log; tail -3 gpurun_out/pytest_$tag.log
for a in 0 1; do
  echo "== staged LRP_STAGE_ASYNC=$a" | tee -a gpurun_out/stage_async_$tag.txt
  LRP_STAGE_ASYNC=$a timeout 300 python bench.py --steps 10 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --variant staged 2>gpurun_out/err_$tag.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])" | tee -a gpurun_out/stage_async_$tag.txt
  LRP_STAGE_ASYNC=$a timeout 600 python tools/bench_configs.py --configs c1t,c3,c4t,c5e,c5p --variants staged --coords fly,table 2>>gpurun_out/err_$tag.txt | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(' ', d['config'], d['coords'], d['us_per_frame'])" | tee -a gpurun_out/stage_async_$tag.txt
done
LRP_STAGE_ASYNC=1 ncu --set full --clock-control none --import-source on -k regex:reproject_staged -s 30 -c 1 -f -o gpurun_out/prof_staged_async_$tag python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --no-sched --e2e-steps 1 --variant staged --coords table > gpurun_out/prof_$tag.log 2>&1; tail -1 gpurun_out/prof_$tag.log
