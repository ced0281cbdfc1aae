#!/bin/bash
# ncu --set full of one launch each (table coordinates = the library default for batches): staged bicubic, nn table kernel,
# gathered bilinear; then the launch list of the default bench command
tag=${1:-r2p}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --no-sched --e2e-steps 1"
ncu --set full --clock-control none --import-source on -k regex:reproject_staged -s 30 -c 1 -f -o gpurun_out/prof_staged_$tag $B > gpurun_out/prof_$tag.log 2>&1; tail -1 gpurun_out/prof_$tag.log
ncu --set full --clock-control none --import-source on -k regex:nn_table -s 30 -c 1 -f -o gpurun_out/prof_nn_$tag $B --interp nn > gpurun_out/prof_$tag.log 2>&1; tail -1 gpurun_out/prof_$tag.log
ncu --set full --clock-control none --import-source on -k regex:reproject_kernel -s 30 -c 1 -f -o gpurun_out/prof_bl_$tag $B --interp bl > gpurun_out/prof_$tag.log 2>&1; tail -1 gpurun_out/prof_$tag.log
ncu --set full --clock-control none --import-source on -k regex:reproject_staged -s 8 -c 1 -f -o gpurun_out/prof_staged_fly_$tag $B --coords fly > gpurun_out/prof_$tag.log 2>&1; tail -1 gpurun_out/prof_$tag.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-sched --e2e-steps 1 > gpurun_out/launches_$tag.log 2>&1
timeout 300 python bench.py --steps 20 --quick --no-cpu-baseline --no-sched --e2e-steps 2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default c2 value', round(d['value'],2), 'fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])"
timeout 300 python bench.py --steps 20 --quick --no-cpu-baseline --no-sched --e2e-steps 2 --interp nn 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nn value', round(d['value'],2), 'frac', round(d['roofline']['frac'],4), 'fly', d['coords_legs']['fly'], 'table', d['coords_legs']['table'])"
ls -la gpurun_out/*$tag*
