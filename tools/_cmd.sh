mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged or channels_rotations" 2>&1 | tail -25 > gpurun_out/racecheck_parity.log; tail -6 gpurun_out/racecheck_parity.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_r.log; tail -3 gpurun_out/pytest_gpu_r.log
