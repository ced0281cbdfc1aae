#!/bin/bash
# after the shared-source buffer cache: scheduler tests, then the N = 1 bench record again
tag=${1:-r2z3}
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_sched.py -q -m gpu 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1_$tag.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'us', round(d['roofline']['us_per_launch'],2), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],2), 'cpu', round(d['cpu_baseline']['value'],4))
for k,v in d['sched'].items(): print(' sched',k,v['frames_per_s'],v['gpix_per_s'],'ceil',v['copy_ceiling_frames_per_s'],v['of_ceiling'],v['run_ms'],v['copy_only_run_ms'])
PY
