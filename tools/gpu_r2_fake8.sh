#!/bin/bash
# scheduler host side under 8 logical devices on 2 physical GPUs: does the engine keep the copy ceiling?
python - <<'PY'
import os, sys, json
os.environ["LRP_FAKE_GPUS"] = os.environ.get("FAKE", "8")
sys.path.insert(0, "."); sys.path.insert(0, "image-lens-reproject_b200/python")
import bench, lrp
lrp.lib()
rot = lrp.rotation_from_degrees(30.0, 20.0, 10.0)
params = lrp.make_params(1, lrp.BICUBIC, rot, None)
n = int(os.environ["LRP_FAKE_GPUS"])
out = bench.run_sched_legs(lrp, n, params, False)
for k, v in out.items():
    print(n, "logical devices", k, json.dumps(v))
PY
