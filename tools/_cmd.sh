mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:png_pack -s 3 -c 1 -f -o gpurun_out/prof_png_pack python tools/bench_encode.py --reps 1 > gpurun_out/prof_png.log 2>&1; tail -2 gpurun_out/prof_png.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:exr_pack -s 3 -c 1 -f -o gpurun_out/prof_exr_pack python tools/bench_encode.py --reps 1 > gpurun_out/prof_exr.log 2>&1; tail -2 gpurun_out/prof_exr.log | cut -c1-200
