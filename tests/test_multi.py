"""N > 1 host logic on CPU: world partition and the end-of-run aggregation of raw-physics_b200/multi.py under a
world_size-2 gloo process group (the GPUs run the same code under NCCL; there is no data-path collective to test)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _multi():
    import __graft_entry__ as ge
    ge.load_package()
    from rawphys_b200 import multi
    return multi


def test_partition_covers_every_world_once():
    multi = _multi()
    for total in (0, 1, 7, 4096, 16384, 4097):
        for g in (1, 2, 3, 4, 8):
            parts = multi.partition(total, g)
            assert len(parts) == g and sum(n for _, n in parts) == total
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(g - 1)) and parts[0][0] == 0
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
    assert multi.owner_of(5, 7, 2) == (1, 1) and multi.owner_of(3, 7, 2) == (0, 3)
    with pytest.raises(IndexError):
        multi.owner_of(7, 7, 2)


def test_single_process_job_is_identity():
    multi = _multi()
    job = multi.Job()
    st = np.random.default_rng(0).normal(size=(3, 4, 21))
    assert job.max_over_ranks(2.5) == 2.5 and np.array_equal(job.sum_over_ranks([1, 2]), [1.0, 2.0])
    assert np.array_equal(job.gather_worlds(st, 3), st)
    assert np.array_equal(job.scatter_worlds(st, 3, (4, 21)), st)


@pytest.mark.parametrize("total", [7, 8])
def test_two_rank_gloo_job(tmp_path, total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multi_worker.py"), str(tmp_path), str(total)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
    res = [json.load(open(tmp_path / ("rank%d.json" % r))) for r in range(2)]
    assert [r["first"] for r in res] == [0, (total + 1) // 2] and sum(r["n"] for r in res) == total
    for r in res:
        assert r["ms"] == 11.0                                  # max over ranks
        assert r["counters"] == [total, 2.0 * total, (total + 1) // 2]  # sums
        assert r["gather_ok"] and r["roundtrip"]
