mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_p.log; tail -3 gpurun_out/pytest_gpu_p.log
python bench.py > gpurun_out/bench_default_p.json 2> gpurun_out/bench_default_p.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_default_p.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['clocks'], d['cpu_baseline']['value'])"
tail -2 gpurun_out/bench_default_p.err
