mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_decode.py tests/test_gpu_sched.py -m gpu -q -x 2>&1 | tail -12
timeout 600 python tests/perf/bench_encode.py 2>&1 | tail -1 > gpurun_out/bench_encode_t.json; python -c "
import json; e=json.load(open('gpurun_out/bench_encode_t.json'))
for k in ('png_decoder','reference_lodepng_decode'): print(k, e.get(k))"
LRP_PNG_UNFILTER_ON_HOST=1 timeout 600 python tests/perf/bench_encode.py 2>&1 | tail -1 | python -c "
import json,sys; e=json.loads(sys.stdin.read()); print('host unfilter:', e.get('png_decoder'))"
timeout 800 python tests/perf/bench_pipeline.py 2>&1 | tail -1 | tee gpurun_out/bench_pipeline_t.json | cut -c1-330
