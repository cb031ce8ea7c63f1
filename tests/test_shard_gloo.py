"""world_size-2 gloo test of the pair sharding + record all-gather (runs on CPU)."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

WORKER = r"""
import os, sys, numpy as np
sys.path.insert(0, os.environ["RF_ROOT"])
import ransac_flow_b200
from ransac_flow_b200 import shard
rank, world, _ = shard.init_from_env("gloo")
n = 7
recs = [shard.pack_record(i, np.eye(3) * (i + 1), nbInlier=10 * i, status=0) for i in shard.my_pairs(n, rank, world)]
allr = shard.gather_records(recs, n, world)
assert allr.shape == (n, shard.RECORD_FLOATS), allr.shape
assert np.array_equal(allr[:, 0], np.arange(n))
assert np.allclose(allr[:, 4], np.arange(n) + 1) and np.allclose(allr[:, 3], 10 * np.arange(n))
print("rank", rank, "ok")
"""


def test_shard_and_gather_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, RF_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_my_pairs_partition():
    sys.path.insert(0, ROOT)
    from ransac_flow_b200 import shard
    for n, w in [(10, 8), (10000, 8), (3, 4), (0, 2)]:
        parts = [shard.my_pairs(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
