mkdir -p gpurun_out
for i in nn bl; do
ncu --set full --clock-control none --import-source on -k regex:reproject_ -s 26 -c 1 -f -o gpurun_out/prof_c2_$i python bench.py --interp $i --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/prof_$i.log 2>&1; tail -1 gpurun_out/prof_$i.log | cut -c1-80
done
ncu --set full --clock-control none --import-source on -k regex:deflate_band -s 1 -c 1 -f -o gpurun_out/prof_deflate python tests/perf/bench_encode.py --reps 1 > gpurun_out/prof_deflate.log 2>&1; tail -1 gpurun_out/prof_deflate.log | cut -c1-80
