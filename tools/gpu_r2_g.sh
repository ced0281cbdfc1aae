#!/bin/bash
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_codec.py tests/test_gpu_decode.py tests/test_gpu_sched.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/pytest_codec_$tag.log; tail -25 gpurun_out/pytest_codec_$tag.log
for lz in 1 0; do LRP_DEFLATE_LZ=$lz timeout 600 python tests/perf/bench_encode.py 2>&1 | tail -3 | cut -c1-1500; done | tee gpurun_out/bench_encode_$tag.txt
