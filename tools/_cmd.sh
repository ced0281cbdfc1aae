mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_codec.py -m gpu -q -x -k "device_deflate_png and (7-33 or 100-109 or 300-500)" 2>&1 | tail -4
timeout 800 python tests/perf/bench_pipeline.py 2>&1 | tail -1 > gpurun_out/bench_pipeline_w.json; python -c "
import json; d=json.load(open('gpurun_out/bench_pipeline_w.json')); print(d['frames_per_s'], d['output_gpix_per_s'])"
