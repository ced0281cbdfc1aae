#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "staged or c2_full or full_size or png_path or exr_path or tiled or pole or c1t or c3_full or lens_matrix or special" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --quick --no-cpu-baseline --no-sched --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])"
timeout 600 python tools/bench_configs.py --configs c1t,c3,c4t,c5e,c5p --variants staged --coords table 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(' ', d['config'], d['coords'], d['us_per_frame'])"
