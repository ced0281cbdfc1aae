mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu_h.log; tail -8 gpurun_out/pytest_gpu_h.log
rm -f gpurun_out/bench_h.jsonl
run() { # env, args
  env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 $2 2>&1 | tail -1 | sed "s/^{/{\"lib\": \"$1\", /" >> gpurun_out/bench_h.jsonl
}
for v in "--interp bc" "--interp bc --coords table"; do run "A=1" "$v"; done
python - <<PY
import json
for l in open("gpurun_out/bench_h.jsonl"):
    try:
        d=json.loads(l); print(d["lib"], d["config"]["interp"], d["config"]["variant"], d["config"]["coords"], round(d["value"],2), "Gpix/s", round(d["roofline"]["us_per_launch"],1), "us", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],2))
    except Exception as e: print(l[:300])
PY
timeout 900 python tools/bench_configs.py --configs c3,c5e,c5p --variants staged 2>&1 | tee gpurun_out/bench_configs_h.jsonl | cut -c1-140
timeout 600 python tools/bench_encode.py 2>&1 | tail -3 | tee gpurun_out/bench_encode_h.json
