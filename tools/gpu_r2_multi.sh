#!/bin/bash
# multi-GPU cycle: the scheduler tests on every physical GPU, then bench.py under torchrun (driver's launch line)
N=${1:-2}; tag=${2:-r2m$N}
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/gpus_$tag.txt; nproc >> gpurun_out/gpus_$tag.txt; nvidia-smi topo -m >> gpurun_out/gpus_$tag.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sched.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_sched_$tag.log; tail -3 gpurun_out/pytest_sched_$tag.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 1500 gpurun_out/bench_$tag.err | tail -5
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "full", round(d["e2e_full_upload"]["value"], 2))
    for k, v in (d.get("sched") or {}).items():
        print(" sched", k, v)
except Exception as e:
    print("unreadable", e)
PY
