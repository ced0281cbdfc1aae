mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_codec.py tests/test_gpu_decode.py -m gpu -q -x -k "not full_size and not 1080 and not ratio" 2>&1 | tail -15 > gpurun_out/sanitizer_codec.log; echo "rc=$?"; tail -8 gpurun_out/sanitizer_codec.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_sched.py tests/test_gpu_roi.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/sanitizer_sched.log; tail -6 gpurun_out/sanitizer_sched.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged or special or lens_matrix" 2>&1 | tail -15 > gpurun_out/sanitizer_parity.log; tail -6 gpurun_out/sanitizer_parity.log
