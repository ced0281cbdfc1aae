mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_default_final.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks']['samples'])"
