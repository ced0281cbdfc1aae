#!/bin/bash
# compute-sanitizer memcheck over what was added late in round 2: both half-warp shapes of the staged kernel, the
# field-of-view mask (table-mode gather kernel with skipped samples, footprints of masked launches), the shared-source cache
cd /root/repo
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "half_warp_shapes" 2>&1 | tail -8 > gpurun_out/r2_sanitizer_memcheck_shapes.log; tail -3 gpurun_out/r2_sanitizer_memcheck_shapes.log
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_fisheye_models.py -m gpu -q -x -k "mask" 2>&1 | tail -8 > gpurun_out/r2_sanitizer_memcheck_mask.log; tail -3 gpurun_out/r2_sanitizer_memcheck_mask.log
