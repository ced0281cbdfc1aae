mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fisheye_models.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_fish.log; tail -15 gpurun_out/pytest_fish.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu_b.log; tail -3 gpurun_out/pytest_gpu_b.log
timeout 900 python tools/bench_configs.py --configs c1,c1t,c4,c4t --coords fly,table --variants staged,gather 2>&1 | tee gpurun_out/bench_configs_b.jsonl | tail -30
