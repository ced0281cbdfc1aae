mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for lib in liblrp.so; do
 for v in "--interp bc" "--interp bc --coords table" "--interp nn" "--interp bl"; do
  LRP_LIB=image-lens-reproject_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 $v 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', d['config']['interp'], d['config']['variant'], d['config']['coords'], round(d['value'],2), 'Gpix/s', round(d['roofline']['us_per_launch'],1), 'us')"
 done
 echo $lib; LRP_LIB=image-lens-reproject_b200/$lib timeout 600 python tools/bench_configs.py --coords fly,table --variants staged 2>&1 | tail -30
done
LRP_LIB=image-lens-reproject_b200/liblrp.so ncu --set full --clock-control none --import-source on -k regex:reproject_ -s 26 -c 1 -f -o gpurun_out/prof_c2_bc_t16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
