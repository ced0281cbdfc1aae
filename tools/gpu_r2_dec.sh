#!/bin/bash
# single-frame decode latency with zlib's inflate and with the built-in one: a 7680x4096 PNG written by the reference's
# lodepng (a real LZ77 stream), a 4K RGBA half EXR (tests/perf/bench_encode.py)
cd /root/repo
for z in 1 0; do
  LRP_INFLATE_ZLIB=$z timeout 200 python tests/perf/bench_encode.py --reps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('zlib=$z png_decoder', round(d['png_decoder']['s'],4), 's  (lodepng decode', round(d['reference_lodepng_decode']['s'],3), 's)  exr_decoder T1', round(d['exr_decoder_T1']['s'],4), 'T16', round(d['exr_decoder_T16']['s'],4))"
done
