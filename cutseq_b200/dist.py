"""Multi-GPU plumbing: one process per GPU, reads sharded as contiguous index ranges.

The trimming path has no exchange step (every read pair is independent), so there is NO collective
on the data path.  ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) is used for exactly three
things: a barrier around timed regions, the max-over-ranks of a measured time, and ONE sum of the
trim statistics (the ``csq_counters`` words) at the end of a run - the analogue of cutadapt merging
the per-worker ``Statistics`` objects (reference run.py:473 / 794 via ``runner.run``).
"""

from __future__ import annotations

import ctypes as C
import os

from . import _abi as A


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, disjoint, order-preserving split of ``range(total)``: rank r gets [lo, hi)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class Group:
    """Thin wrapper so that the same code runs with world_size 1 (no process group at all)."""

    def __init__(self, backend: str | None = None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            kwargs = {}
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self.device = torch.device("cuda", self.local_rank)
                kwargs["device_id"] = self.device
            else:
                self.device = torch.device("cpu")
            if not dist.is_initialized():
                dist.init_process_group(backend, rank=self.rank, world_size=self.world, **kwargs)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    # A barrier on the GPUs is a kernel that spins until the last rank arrives: ranks that only wait for rank 0's host-side
    # work (the whole-file leg drives every GPU of the job from rank 0) must not do that - their spinning kernel would
    # time-slice with the file run on their GPU.  These two go through the rendezvous store (a socket) instead.
    def host_signal(self, key: str):
        if self.dist is not None:
            self.dist.distributed_c10d._get_default_store().set(key, "1")

    def host_wait(self, key: str, timeout_s: float = 3600.0):
        if self.dist is not None:
            import datetime

            self.dist.distributed_c10d._get_default_store().wait([key], datetime.timedelta(seconds=timeout_s))

    def max(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_counters(self, counters: A.csq_counters) -> A.csq_counters:
        """The one reduction of trim statistics: element-wise sum of the csq_counters words."""
        if self.dist is None:
            return counters
        import torch

        n = C.sizeof(A.csq_counters) // 8
        words = (C.c_uint64 * n).from_buffer_copy(bytes(counters))
        t = torch.tensor(list(words), dtype=torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        out = A.csq_counters.from_buffer_copy((C.c_uint64 * n)(*[int(v) for v in t.tolist()]))
        return out

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
            self.dist = None
