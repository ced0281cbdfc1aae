mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_decode.py -m gpu -q -x 2>&1 | tail -8
