"""One rank of the world_size-2 gloo job of tests/test_multi.py (CPU only): exercises raw-physics_b200/multi.py -- the
partition of worlds over ranks and the end-of-run aggregation -- exactly as bench.py drives it under NCCL."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.load_package()
    from rawphys_b200 import multi
    out_dir, total = sys.argv[1], int(sys.argv[2])
    dist.init_process_group("gloo")
    job = multi.Job(dist, "cpu")
    first, n = job.my_worlds(total)
    NB, S = 5, 21
    # every (world, body, field) gets a value that names it, so misplaced blocks are visible
    w = np.arange(first, first + n, dtype=np.float64)[:, None, None]
    state = w * 1000.0 + np.arange(NB)[None, :, None] * 30.0 + np.arange(S)[None, None, :]
    status = (np.arange(first, first + n) % 3).astype(np.int32)
    res = {}
    res["ms"] = job.max_over_ranks(10.0 + job.rank)
    res["counters"] = job.sum_over_ranks([n, 2.0 * n, first]).tolist()
    full = job.gather_worlds(state, total)
    st = job.gather_worlds(status, total)
    ck = os.path.join(out_dir, "ck.npz")
    job.save_checkpoint(ck, state, total, frame=17)
    back, frame = job.load_checkpoint(ck, total, (NB, S))
    res["roundtrip"] = bool(np.array_equal(back, state)) and frame == 17
    if job.rank == 0:
        want = np.arange(total, dtype=np.float64)[:, None, None] * 1000.0 + np.arange(NB)[None, :, None] * 30.0 + np.arange(S)[None, None, :]
        res["gather_ok"] = bool(np.array_equal(full, want)) and bool(np.array_equal(st, np.arange(total) % 3))
    else:
        res["gather_ok"] = full is None and st is None
    res["first"], res["n"] = first, n
    import json
    with open(os.path.join(out_dir, "rank%d.json" % job.rank), "w") as f:
        json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
