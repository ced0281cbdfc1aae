mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_codec.py tests/test_gpu_decode.py tests/test_gpu_sched.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python tests/perf/bench_encode.py 2>&1 | tail -1 > gpurun_out/bench_encode_u.json; python -c "
import json; e=json.load(open('gpurun_out/bench_encode_u.json'))
for k in ('png_encoder_device','exr_encoder_device'): print(k, e.get(k))"
