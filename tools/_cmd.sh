mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default_s.json 2> gpurun_out/bench_default_s.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_default_s.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['remap_table_variant'], d['e2e']['value'], d['clocks'])"
tail -2 gpurun_out/bench_default_s.err
