#!/bin/bash
# round 2: the device inflate on a lowest-priority stream (library streams at the highest): decode tests, EXR pipeline A/B
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_sched.py -x -q -m gpu 2>&1 | tail -2
for t in 16 32; do
  for pr in 0 1; do
    LRP_INFLATE_PRIORITY=$pr timeout 600 python tests/perf/bench_pipeline.py --exr --frames 128 --threads $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('exr pipeline threads $t priority-split $pr: host inflate', d['host_inflate']['frames_per_s'], 'device inflate', d['device_inflate']['frames_per_s'])"
  done
done
