"""Worlds across the GPUs of one box: one process per GPU, every rank owns a full batch of its own worlds,
NO collective on the data path -- worlds never interact (no shared state; hulls are read-only and replicated). The
process group (NCCL on the GPUs, gloo in the CPU tests) carries only what is aggregated at the END of a run: the timing
(max over ranks), the work counters (sum), per-world status words and state checkpoints (gather to rank 0 / scatter back).

Nothing here touches a batch: the functions take and return numpy arrays / numbers, so the same code runs under
`torchrun` on 8 B200s (bench.py) and in the world_size-2 gloo tests (tests/test_multi.py).
"""
import numpy as np


def partition(total_worlds, world_size):
    """Contiguous blocks, sizes differing by at most one: [(first_world, n_worlds)] per rank. World w of the job is world
    w - first of the rank that owns it; G-GPU results equal 1-GPU results per world because worlds are independent."""
    if total_worlds < 0 or world_size < 1:
        raise ValueError("partition(%r, %r)" % (total_worlds, world_size))
    base, extra = divmod(total_worlds, world_size)
    out, first = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((first, n))
        first += n
    return out


def owner_of(world, total_worlds, world_size):
    """(rank, local index) of a job-wide world id under partition()."""
    for r, (first, n) in enumerate(partition(total_worlds, world_size)):
        if first <= world < first + n:
            return r, world - first
    raise IndexError(world)


class Job:
    """The ranks of one run. `dist` is torch.distributed (initialised) or None for a single process; `device` is where
    the small reduction tensors live ("cuda" under NCCL, "cpu" under gloo)."""

    def __init__(self, dist=None, device="cpu"):
        self.dist = dist
        self.device = device
        self.rank = dist.get_rank() if dist is not None else 0
        self.world_size = dist.get_world_size() if dist is not None else 1

    def my_worlds(self, total_worlds):
        return partition(total_worlds, self.world_size)[self.rank]

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x):
        """Device time of a multi-GPU run is the slowest rank's."""
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, values):
        """Work counters (pair tests, EPA runs, contacts ...) of the whole job."""
        v = np.asarray(values, dtype=np.float64)
        if self.dist is None:
            return v.copy()
        import torch
        t = torch.from_numpy(v.copy()).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def gather_worlds(self, local, total_worlds):
        """Per-world arrays ([n_local, ...]: state records, status words) of every rank -> one [total_worlds, ...] array in
        job-wide world order on rank 0 (None elsewhere). Ranks may own different numbers of worlds."""
        local = np.ascontiguousarray(local)
        first, n = self.my_worlds(total_worlds)
        if local.shape[0] != n:
            raise ValueError("rank %d owns %d worlds, got %d" % (self.rank, n, local.shape[0]))
        if self.dist is None:
            return local.copy()
        import torch
        parts = partition(total_worlds, self.world_size)
        most = max(p[1] for p in parts)
        pad = np.zeros((most,) + local.shape[1:], dtype=local.dtype)
        pad[:n] = local
        mine = torch.from_numpy(pad).to(self.device)
        bufs = [torch.empty_like(mine) for _ in range(self.world_size)] if self.rank == 0 else None
        self.dist.gather(mine, bufs, dst=0)
        if self.rank != 0:
            return None
        return np.concatenate([bufs[r].cpu().numpy()[:parts[r][1]] for r in range(self.world_size)], axis=0)

    def scatter_worlds(self, full, total_worlds, tail_shape, dtype=np.float64):
        """Inverse of gather_worlds: rank 0 holds [total_worlds, ...]; every rank gets its own block."""
        first, n = self.my_worlds(total_worlds)
        if self.dist is None:
            return np.ascontiguousarray(full[first:first + n]).astype(dtype, copy=True)
        import torch
        parts = partition(total_worlds, self.world_size)
        most = max(p[1] for p in parts)
        out = torch.empty((most,) + tuple(tail_shape), dtype=torch.from_numpy(np.zeros(0, dtype=dtype)).dtype, device=self.device)
        chunks = None
        if self.rank == 0:
            full = np.ascontiguousarray(full, dtype=dtype)
            if full.shape != (total_worlds,) + tuple(tail_shape):
                raise ValueError("scatter_worlds: %r != %r" % (full.shape, (total_worlds,) + tuple(tail_shape)))
            chunks = []
            for (f, k) in parts:
                pad = np.zeros((most,) + tuple(tail_shape), dtype=dtype)
                pad[:k] = full[f:f + k]
                chunks.append(torch.from_numpy(pad).to(self.device))
        self.dist.scatter(out, chunks, src=0)
        return out.cpu().numpy()[:n].copy()

    # ---- checkpoints: the whole job's per-world state in one file, written and read by rank 0
    def save_checkpoint(self, path, local_state, total_worlds, frame, meta=None):
        full = self.gather_worlds(local_state, total_worlds)
        if self.rank == 0:
            np.savez(path, state=full, frame=np.int64(frame), total_worlds=np.int64(total_worlds), **(meta or {}))
        self.barrier()

    def load_checkpoint(self, path, total_worlds, tail_shape):
        """-> (this rank's [n_local, ...] state block, frame)."""
        full, frame = None, 0
        if self.rank == 0:
            z = np.load(path)
            if int(z["total_worlds"]) != total_worlds:
                raise ValueError("checkpoint holds %d worlds, the job has %d" % (int(z["total_worlds"]), total_worlds))
            full, frame = z["state"], int(z["frame"])
        frame = int(self.sum_over_ranks([frame if self.rank == 0 else 0])[0])
        return self.scatter_worlds(full, total_worlds, tail_shape), frame
