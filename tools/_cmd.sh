mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "not 1080" 2>&1 | tail -6 > gpurun_out/sanitizer_decode.log; tail -4 gpurun_out/sanitizer_decode.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 99 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "pillow or written_by_the_reference" 2>&1 | tail -8 > gpurun_out/racecheck_decode.log; tail -5 gpurun_out/racecheck_decode.log
