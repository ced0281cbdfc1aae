#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python tests/perf/bench_pipeline.py --frames 96 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pipeline png threads', d['threads'], 'fps', round(d['frames_per_s'],2))"
