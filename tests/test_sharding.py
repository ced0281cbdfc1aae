"""The multi-GPU form of the path: a partition of the batch over ranks with no
data-path collective.  Host logic only — runs on CPU: shard arithmetic, and a world_size-2 `gloo` run of the
exact reduction bench.py performs (max of the device times, sum of the units, whole-job rate)."""
import json
import os
import subprocess
import sys

import pytest

import oracle_lib as ol

sys.path.insert(0, os.path.join(ol.ROOT, "image-lens-reproject_b200", "python"))
from lrp import sharding  # noqa: E402  (pure python: does not load liblrp.so)


@pytest.mark.parametrize("n,world", [(1024, 8), (6, 4), (6, 8), (0, 3), (7, 1), (5, 2), (1000, 7)])
def test_shard_range_is_a_balanced_partition(n, world):
    parts = [sharding.shard_range(n, r, world) for r in range(world)]
    flat = [i for p in parts for i in p]
    assert flat == list(range(n))  # disjoint, ordered, complete
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1
    assert sizes == sorted(sizes, reverse=True)


def test_shard_range_named_configs():
    assert [len(sharding.shard_range(1024, r, 8)) for r in range(8)] == [128] * 8  # c4
    assert [len(sharding.shard_range(6, r, 4)) for r in range(4)] == [2, 2, 1, 1]  # c5 views
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_weak_batch_union_is_world_independent():
    for world in (1, 2, 4, 8):
        frames = sorted(i for r in range(world) for i in sharding.weak_batch(8, r, world))
        assert frames == list(range(8 * world))


def test_single_process_reductions_are_identity():
    assert sharding.max_over_ranks([1.5, 2.0]) == [1.5, 2.0]
    assert sharding.sum_over_ranks([3, 4]) == [3.0, 4.0]
    assert sharding.whole_job_rate(100.0, 4.0) == 25.0


_WORKER = r"""
import json, os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from lrp import sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
frames = list(sharding.weak_batch(8, rank, world))
views = list(sharding.shard_range(6, rank, world))
ms = 10.0 + 5.0 * rank                     # rank 1 is the slow one
dist.barrier()
t_max, e_max = sharding.max_over_ranks([ms, 2 * ms], dist)
units, = sharding.sum_over_ranks([len(frames) * 100], dist)
rate = sharding.whole_job_rate(units, t_max * 1e-3)
gathered = [None] * world
dist.all_gather_object(gathered, {"rank": rank, "frames": frames, "views": views})
dist.barrier()
if rank == 0:
    print(json.dumps({"t_max": t_max, "e_max": e_max, "units": units, "rate": rate, "parts": gathered}))
dist.destroy_process_group()
"""


def test_world_size_2_gloo_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    pkg = os.path.join(ol.ROOT, "image-lens-reproject_b200", "python")
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script), pkg]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["t_max"] == 15.0 and d["e_max"] == 30.0          # the slowest rank defines the time
    assert d["units"] == 1600.0                               # 2 ranks x 8 frames x 100
    assert abs(d["rate"] - 1600.0 / 0.015) < 1e-6
    parts = sorted(d["parts"], key=lambda p: p["rank"])
    assert parts[0]["frames"] == list(range(0, 8)) and parts[1]["frames"] == list(range(8, 16))
    assert parts[0]["views"] == [0, 1, 2] and parts[1]["views"] == [3, 4, 5]


def test_reference_arm_runs_on_rank_0_only():
    """under torchrun the CPU reference arm is rank 0's job; the other ranks exit 0 without output"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ol.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
