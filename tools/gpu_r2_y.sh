#!/bin/bash
# round 2, step y: single steps of a split tile as 8 x 4 blocks (LRP_ST_SPLIT_BLOCKS) — parity, then A/B on every config
cd /root/repo
P=/root/repo/image-lens-reproject_b200
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_roi.py -x -q -m gpu 2>&1 | tail -4
for lib in liblrp_nosplit.so liblrp.so; do
  LRP_LIB=$P/$lib timeout 300 python bench.py --steps 20 --quick --no-cpu-baseline --no-sched --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib c2 fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])"
  for cm in table fly; do
  LRP_LIB=$P/$lib timeout 600 python tools/bench_configs.py --configs c1t,c3,c4t,c5e,c5p --variants staged --coords $cm 2>/dev/null | python -c "
import json,sys
print('   $cm', ' '.join('%s %s' % (json.loads(l)['config'], json.loads(l)['us_per_frame']) for l in sys.stdin))"
  done
done
