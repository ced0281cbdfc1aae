#!/bin/bash
# ncu --set full of one launch each: tiled / staged bicubic in table mode, the nn table kernel
tag=${1:-r2p}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --no-sched --e2e-steps 1"
ncu --set full --clock-control none --import-source on -k regex:reproject_tiled -s 30 -c 1 -f -o gpurun_out/prof_tiled_$tag $B --variant tiled --coords table > gpurun_out/prof_$tag.log 2>&1; tail -2 gpurun_out/prof_$tag.log
ncu --set full --clock-control none --import-source on -k regex:reproject_staged -s 30 -c 1 -f -o gpurun_out/prof_staged_$tag $B --variant staged --coords table > gpurun_out/prof_$tag.log 2>&1; tail -2 gpurun_out/prof_$tag.log
ncu --set full --clock-control none --import-source on -k regex:nn_table -s 30 -c 1 -f -o gpurun_out/prof_nn_$tag $B --interp nn --coords table > gpurun_out/prof_$tag.log 2>&1; tail -2 gpurun_out/prof_$tag.log
ls -la gpurun_out/*.ncu-rep
