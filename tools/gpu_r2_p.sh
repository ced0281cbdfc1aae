#!/bin/bash
# round 2: half-warp shape per launch (rows / 4 x 4 blocks for pole-crossing views): parity, then the configs A/B
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for b in 0 auto; do
  if [ $b = auto ]; then unset LRP_ST_BLOCKS; else export LRP_ST_BLOCKS=$b; fi
  for cm in table fly; do
  timeout 600 python tools/bench_configs.py --configs c1t,c5e,c5p --variants auto --coords $cm 2>/dev/null | python -c "
import json,sys
print('blocks=$b $cm', ' '.join('%s %s' % (json.loads(l)['config'], json.loads(l)['us_per_frame']) for l in sys.stdin))"
  done
done
unset LRP_ST_BLOCKS
timeout 300 python bench.py --steps 20 --quick --no-cpu-baseline --no-sched --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])"
