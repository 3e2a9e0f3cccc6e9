"""CPU: the N>1 host logic with world_size-2 gloo (sharding at TB granularity, counter all-reduce)."""
import os
import subprocess
import sys

import pytest

from neoradium_b200 import dist as nd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for n in (1, 7, 64, 65, 1000):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                lo, hi = nd.shard_range(n, r, w)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [nd.shard_range(n, r, w)[1] - nd.shard_range(n, r, w)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


WORKER = r"""
import os, sys
sys.path.insert(0, '@ROOT@')
import torch, torch.distributed as dist
from neoradium_b200 import dist as nd
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = nd.shard_range(101, r, w)
c = torch.zeros(8, dtype=torch.int64)
c[0] = (hi - lo) * 16; c[2] = hi - lo; c[3] = r + 1; c[5] = (hi - lo) * 16 * 8
nd.reduce_counters(c)
d = nd.counters_dict(c)
assert d["txBlocks"] == 101 and d["codeBlocks"] == 1616 and d["tbCrcFail"] == 3 and d["meanIterations"] == 8.0, d
open(os.path.join(os.environ["NR_OUT"], "rank%d.ok" % r), "w").write("ok")
dist.destroy_process_group()
"""


def test_counter_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER.replace('@ROOT@', ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", NR_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()
