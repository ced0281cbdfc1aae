mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sched.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu_q.log; tail -25 gpurun_out/pytest_gpu_q.log
