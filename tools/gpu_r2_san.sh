#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels (nearest table / byte map, tiled, supersampling, LZ deflate, engine)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "nearest_byte_map_u8 or nearest_texel_copy or coords_modes or remap_cache or tiled_kernel or supersampling_staged" 2>&1 | tail -12 > gpurun_out/r2_sanitizer_memcheck_round2.log; tail -4 gpurun_out/r2_sanitizer_memcheck_round2.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_codec.py -m gpu -q -x -k "finds_the_matches or valid_zlib or rendered_style" 2>&1 | tail -12 > gpurun_out/r2_sanitizer_memcheck_deflate.log; tail -4 gpurun_out/r2_sanitizer_memcheck_deflate.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_sched.py -m gpu -q -x -k "engine or levels or copy_only or shared_source or wait_on or async_submit or wide_then_tall" 2>&1 | tail -12 > gpurun_out/r2_sanitizer_memcheck_sched.log; tail -4 gpurun_out/r2_sanitizer_memcheck_sched.log
