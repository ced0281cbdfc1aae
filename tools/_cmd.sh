mkdir -p gpurun_out
timeout 800 python tools/bench_pipeline.py 2>&1 | tail -2 | tee gpurun_out/bench_pipeline.json | cut -c1-1500
