"""bench.py's reference arm runs without a GPU (the reference's own CPU code from oracle/_ref, else the oracle port):
its JSON line must keep the driver's contract.  The GPU arm must refuse to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import oracle_lib as ol

ROOT = ol.ROOT


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "output_gpix_per_s" and line["unit"] == "Gpix/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["config"]["workload"].startswith("c2:")
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
