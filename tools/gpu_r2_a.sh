#!/bin/bash
# round 2, first GPU cycle: new parity tests first (fast feedback), then the whole suite, then the bench
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$tag.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_sched.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_new_$tag.log; tail -5 gpurun_out/pytest_new_$tag.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu_$tag.log; tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 10 --quick --no-cpu-baseline > gpurun_out/bench_quick_$tag.json 2> gpurun_out/bench_quick_$tag.err; tail -c 600 gpurun_out/bench_quick_$tag.err
( time timeout 900 python bench.py ) > gpurun_out/bench_full_$tag.json 2> gpurun_out/bench_full_$tag.err; tail -c 400 gpurun_out/bench_full_$tag.err
python - <<PY
import json
for f in ("gpurun_out/bench_quick_$tag.json", "gpurun_out/bench_full_$tag.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 2), "us", round(d["roofline"]["us_per_launch"], 1), "frac", round(d["roofline"]["frac"], 4),
              "legs", d["coords_legs"]["fly"]["us_per_launch"], d["coords_legs"]["table"]["us_per_launch"], "e2e", round(d["e2e"]["value"], 2),
              "full", round(d["e2e_full_upload"]["value"], 2))
        print(" interp", d.get("interp_legs"))
        print(" configs", d.get("configs"))
        print(" sched", d.get("sched"))
    except Exception as e:
        print(f, "unreadable", e)
PY
