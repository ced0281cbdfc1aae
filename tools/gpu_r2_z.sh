#!/bin/bash
# round 2, final single-GPU cycle: smoke, the whole GPU suite, pipeline worker count, the bench record of both arms
tag=${1:-r2z2}
cd /root/repo
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
for t in 16 24 32; do
  timeout 600 python tests/perf/bench_pipeline.py --frames 96 --threads $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pipeline png threads', d['threads'], 'fps', round(d['frames_per_s'],2))"
done
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
timeout 1200 python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1_$tag.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'us', round(d['roofline']['us_per_launch'],2), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],2), 'cpu', round(d['cpu_baseline']['value'],4))
for k,v in d['sched'].items(): print(' sched',k,v['frames_per_s'],v['gpix_per_s'],'ceil',v['copy_ceiling_frames_per_s'],v['of_ceiling'],v['run_ms'],v['copy_only_run_ms'])
r=json.loads(open('gpurun_out/bench_ref_$tag.json').read().strip().splitlines()[-1]); print('ref', r['value'])
PY
