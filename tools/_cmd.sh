mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu_k.log; tail -3 gpurun_out/pytest_gpu_k.log
python bench.py > gpurun_out/bench_default_k.json 2> gpurun_out/bench_default_k.err; tail -c 600 gpurun_out/bench_default_k.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_k.json 2>&1; tail -c 300 gpurun_out/bench_reference_k.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reproject_ -s 26 -c 1 -f -o gpurun_out/prof_c2_bc_k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/prof_k.log 2>&1; tail -1 gpurun_out/prof_k.log | cut -c1-100
timeout 600 python tools/bench_encode.py 2>&1 | tail -1 > gpurun_out/bench_encode_k.json; cut -c1-300 gpurun_out/bench_encode_k.json
