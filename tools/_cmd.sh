mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu_a.log; tail -3 gpurun_out/pytest_gpu_a.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default_a.json; cut -c1-300 gpurun_out/bench_default_a.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference_a.json; cut -c1-400 gpurun_out/bench_reference_a.json
timeout 900 python tools/bench_configs.py --coords fly,table --variants staged,gather 2>&1 | tee gpurun_out/bench_configs_a.jsonl | tail -30
