mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu_e.log; tail -3 gpurun_out/pytest_gpu_e.log
rm -f gpurun_out/bench_e.jsonl
run() { # env, args
  env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 $2 2>&1 | tail -1 | sed "s/^{/{\"lib\": \"$1\", /" >> gpurun_out/bench_e.jsonl
}
for v in "--interp bc" "--interp bl" "--interp nn" "--interp bc --coords table" "--interp bc --variant gather" "--interp bl --variant gather" "--interp nn --variant gather"; do run "A=1" "$v"; done
python - <<PY
import json
for l in open("gpurun_out/bench_e.jsonl"):
    try:
        d=json.loads(l); print(d["lib"], d["config"]["interp"], d["config"]["variant"], d["config"]["coords"], round(d["value"],2), "Gpix/s", round(d["roofline"]["us_per_launch"],1), "us", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],2))
    except Exception as e: print(l[:300])
PY
timeout 900 python tools/bench_configs.py --configs c1t,c3,c4t,c5e,c5p --variants staged,gather 2>&1 | tee gpurun_out/bench_configs_e.jsonl | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:reproject_ -s 26 -c 1 -f -o gpurun_out/prof_c2_bc_e python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/prof_e.log 2>&1; tail -1 gpurun_out/prof_e.log | cut -c1-200
