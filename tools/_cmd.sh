mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for v in "--interp bc" "--interp nn" "--interp bl"; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 $v 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['interp'], round(d['value'],2), round(d['roofline']['us_per_launch'],1), round(d['remap_table_variant']['us_per_launch'],1))"; done
timeout 600 python tools/bench_configs.py --configs c3,c4t,c5e --variants staged 2>&1 | cut -c1-120
