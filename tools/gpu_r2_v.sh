#!/bin/bash
# round 2, step v: field-of-view mask + fast host inflate on the GPU box: tests, then the file -> file pipeline A/B
tag=${1:-r2v}
mkdir -p gpurun_out
cd /root/repo
timeout 900 python -m pytest tests/test_fisheye_models.py tests/test_gpu_decode.py tests/test_gpu_codec.py -x -q -m gpu > gpurun_out/pytest_$tag.log 2>&1
tail -5 gpurun_out/pytest_$tag.log
for z in 0 1; do
  LRP_INFLATE_ZLIB=$z timeout 600 python tests/perf/bench_pipeline.py --frames 64 > gpurun_out/pipeline_png_zlib${z}_$tag.json 2> gpurun_out/pipeline_png_zlib${z}_$tag.err
  tail -c 1500 gpurun_out/pipeline_png_zlib${z}_$tag.json
  LRP_INFLATE_ZLIB=$z timeout 600 python tests/perf/bench_pipeline.py --exr --frames 128 > gpurun_out/pipeline_exr_zlib${z}_$tag.json 2> gpurun_out/pipeline_exr_zlib${z}_$tag.err
  tail -c 1500 gpurun_out/pipeline_exr_zlib${z}_$tag.json
done
nproc; grep -m1 "model name" /proc/cpuinfo
