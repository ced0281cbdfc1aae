#!/bin/bash
# One GPU iteration: parity tests, bench variants, ncu full capture of the hot kernel.
# usage (under gpurun): bash tools/gpu_cycle.sh <tag> [pytest-args]
tag=${1:-x}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x "$@" 2>&1 | tail -25 > gpurun_out/pytest_gpu_$tag.log; tail -3 gpurun_out/pytest_gpu_$tag.log
rm -f gpurun_out/bench_variants_$tag.jsonl
for v in "--interp bc" "--interp bl" "--interp nn" "--interp bc --coords table" "--interp bc --variant gather" "--interp bl --variant gather" "--interp nn --variant gather" "--interp bc --variant gather --coords table"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 $v 2>&1 | tail -1 >> gpurun_out/bench_variants_$tag.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/bench_variants_$tag.jsonl"):
    try:
        d=json.loads(l); print(d["config"]["interp"], d["config"]["variant"], d["config"]["coords"], round(d["value"],2), "Gpix/s", round(d["roofline"]["us_per_launch"],1), "us", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],2))
    except Exception as e: print(l[:300])
PY
ncu --set full --clock-control none --import-source on -k regex:reproject_ -s 26 -c 1 -f -o gpurun_out/prof_c2_bc_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/prof_$tag.log 2>&1; tail -1 gpurun_out/prof_$tag.log
