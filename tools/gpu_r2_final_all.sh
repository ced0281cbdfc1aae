#!/bin/bash
# end-of-round record on one GPU: smoke, whole GPU suite, reference arm, full bench line, encode bench
tag=${1:-r2final}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_$tag.log; tail -3 gpurun_out/pytest_gpu_$tag.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
( time python bench.py ) > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err; tail -4 gpurun_out/bench_n1_$tag.err
timeout 600 python tests/perf/bench_encode.py 2>&1 | tail -1 > gpurun_out/bench_encode_$tag.json
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n1_$tag.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "us", round(d["roofline"]["us_per_launch"], 1), "frac", round(d["roofline"]["frac"], 4), "legs", d["coords_legs"]["fly"], d["coords_legs"]["table"])
print("e2e", round(d["e2e"]["value"], 2), "full", round(d["e2e_full_upload"]["value"], 2), "cpu", d["cpu_baseline"])
for k in ("interp_legs", "configs", "supersampling", "sched"):
    print(k, json.dumps(d.get(k)))
r = json.loads(open("gpurun_out/bench_ref_$tag.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"]["cores"], "same config", r["config"] == d["config"])
e = json.loads(open("gpurun_out/bench_encode_$tag.json").read())
print("encode", {k: e[k] for k in ("png_encoder_device", "exr_encoder_device")})
PY
