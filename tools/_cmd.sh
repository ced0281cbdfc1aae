mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_codec.py -m gpu -q -x -k "distribution" 2>&1 | tail -25 > gpurun_out/pytest_gpu_o.log; tail -25 gpurun_out/pytest_gpu_o.log
