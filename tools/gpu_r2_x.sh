#!/bin/bash
# round 2, step x: lane-replicated quantiser thresholds (LRP_THR_REPL) and 4 x 4 half-warp blocks (LRP_ST_MAP): parity of
# the combined build, then c2 / c1t / c5e / c5p timings of the four builds at the record pitches the bank model names
cd /root/repo
P=/root/repo/image-lens-reproject_b200
LRP_LIB=$P/liblrp_both.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_roi.py -x -q -m gpu 2>&1 | tail -4
run() { # lib pad
  LRP_REC_PAD=$2 LRP_LIB=$P/$1 timeout 300 python bench.py --steps 20 --quick --no-cpu-baseline --no-sched --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 pad $2 c2 fly', d['coords_legs']['fly']['us_per_launch'], 'table', d['coords_legs']['table']['us_per_launch'])"
}
cfg() {
  LRP_REC_PAD=$2 LRP_LIB=$P/$1 timeout 600 python tools/bench_configs.py --configs c1t,c5e,c5p --variants staged --coords table 2>/dev/null | python -c "
import json,sys
print('   ', ' '.join('%s %s' % (json.loads(l)['config'], json.loads(l)['us_per_frame']) for l in sys.stdin))"
}
run liblrp_base.so 17; cfg liblrp_base.so 17
run liblrp.so 17; cfg liblrp.so 17
for pad in 20 28 22 17; do run liblrp_map.so $pad; done
cfg liblrp_map.so 20
for pad in 20 28 22; do run liblrp_both.so $pad; done
cfg liblrp_both.so 20
