"""N>1 host logic on CPU: world_size-2 gloo process group (sharding, max-over-ranks, counter reduction)."""

import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

from cutseq_b200 import dist
from tests import helpers


def test_shard_ranges_partition_the_workload():
    for total in (0, 1, 7, 10, 200_000_000):
        for world in (1, 2, 3, 4, 8):
            pieces = [dist.shard_range(total, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == total
            for (a, b), (c, d) in zip(pieces, pieces[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dist.shard_range(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, helpers.ROOT)
    from cutseq_b200 import _abi as A
    from cutseq_b200 import dist as d
    from cutseq_b200 import native
    from oracle import oracle

    g = d.Group(backend="gloo")
    # each rank trims its contiguous shard of one synthetic workload on the CPU oracle (no GPU here);
    # the reduction of the statistics must equal the single-process result
    total = 600
    lo, hi = d.shard_range(total, rank, world)
    prog = helpers.program_for(["-A", "TAKARAV3", "--trim-polyA"], 2)
    batch = native.synth_batch(2, hi - lo, first_index=lo, buffer=rank)
    out = oracle.run_batch(prog, batch, want_matches=False)
    summed = g.sum_counters(out["counters"])
    slowest = g.max(float(rank + 1))
    g.barrier()
    # host-side wait (no collective): rank 1 waits until rank 0 has "finished its host work"
    if rank == 0:
        import time

        time.sleep(0.3)
        g.host_signal("rank0_done")
    else:
        g.host_wait("rank0_done", timeout_s=60)
    q.put((rank, lo, hi, summed.n, summed.written, summed.too_short, list(summed.quality_trimmed_bp), slowest,
           out["text"][0][0][:64]))
    g.close()


def test_two_rank_gloo_reduction_matches_single_process():
    from cutseq_b200 import native
    from oracle import oracle

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prog = helpers.program_for(["-A", "TAKARAV3", "--trim-polyA"], 2)
    whole = oracle.run_batch(prog, native.synth_batch(2, 600, first_index=0, buffer=0), want_matches=False)
    c = whole["counters"]
    for rank, lo, hi, n, written, short, qt, slowest, head in results:
        assert (n, written, short, qt) == (c.n, c.written, c.too_short, list(c.quality_trimmed_bp))
        assert slowest == 2.0
    assert results[0][1:3] == (0, 300) and results[1][1:3] == (300, 600)
    assert results[0][8] == whole["text"][0][0][:64]  # rank 0's shard starts the ordered output
