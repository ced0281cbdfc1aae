#!/bin/bash
# the last single-GPU cycle of round 2: smoke, the whole GPU suite, both bench arms
tag=${1:-r2end}
cd /root/repo
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -2 gpurun_out/pytest_gpu_$tag.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
timeout 1200 python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1_$tag.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'us', round(d['roofline']['us_per_launch'],2), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],2), 'cpu', round(d['cpu_baseline']['value'],4))
print(' configs', {k: (v['us'], v['fly_us']) for k, v in d['configs'].items()})
for k,v in d['sched'].items(): print(' sched',k,v['frames_per_s'],v['gpix_per_s'],'ceil',v['copy_ceiling_frames_per_s'],v['of_ceiling'])
r=json.loads(open('gpurun_out/bench_ref_$tag.json').read().strip().splitlines()[-1]); print('ref', r['value'])
PY
