mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for lib in liblrp.so liblrp_w24.so liblrp_w32.so; do
 for v in "--interp bc" "--interp nn" "--interp bc --coords table"; do
  LRP_LIB=image-lens-reproject_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 $v 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', d['config']['interp'], d['config']['variant'], d['config']['coords'], round(d['value'],2), 'Gpix/s', round(d['roofline']['us_per_launch'],1), 'us')"
 done
done
