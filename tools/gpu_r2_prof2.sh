#!/bin/bash
# ncu --set full of the c5 pole view's staged kernel (table coordinates) in both half-warp shapes: rows forced, blocks (default)
cd /root/repo
mkdir -p gpurun_out
B="python tools/bench_configs.py --configs c5p --variants auto --coords table"
LRP_ST_BLOCKS=0 ncu --set full --clock-control none --import-source on -k regex:reproject_staged -s 4 -c 1 -f -o gpurun_out/prof_c5p_rows $B > gpurun_out/prof_c5p_rows.log 2>&1; tail -1 gpurun_out/prof_c5p_rows.log
ncu --set full --clock-control none --import-source on -k regex:reproject_staged -s 4 -c 1 -f -o gpurun_out/prof_c5p_blocks $B > gpurun_out/prof_c5p_blocks.log 2>&1; tail -1 gpurun_out/prof_c5p_blocks.log
ls -la gpurun_out/prof_c5p_*
