#!/usr/bin/env python
"""Host-side probes for the file driver (run on the GPU box): the text reader alone, csq_run_files without
outputs, and the whole thing; prints pairs/s for each so the slow stage can be named."""
import ctypes as C
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from cutseq_b200 import _abi as A  # noqa: E402
from cutseq_b200 import native  # noqa: E402


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    tmp = tempfile.mkdtemp()
    batch = native.synth_batch(2, pairs, first_index=0, buffer=0)
    ins = [os.path.join(tmp, f"in_R{m}.fq") for m in (1, 2)]
    for m, p in enumerate(ins):
        with open(p, "wb") as f:
            f.write(memoryview(native.format_fastq(batch, m)))
    L = native.lib()
    native.device_count()  # CUDA context up: the reader's buffers are pinned
    for rep in range(2):
        t0 = time.time()
        h = C.c_void_p()
        native.check(L.csq_text_reader_open(ins[0].encode(), ins[1].encode(), C.byref(h)))
        tot, i, ts = 0, 0, []
        while True:
            bt = A.csq_batch_text()
            t1 = time.time()
            native.check(L.csq_text_reader_next(h, i % 4, 1 << 17, C.byref(bt)))
            ts.append(time.time() - t1)
            if bt.n_reads == 0:
                break
            tot += bt.n_reads
            i += 1
        L.csq_text_reader_close(h)
        dt = time.time() - t0
        print(f"text reader alone: {tot / dt / 1e6:.2f} M pairs/s; ms per batch: {[round(x * 1e3) for x in ts[:12]]}", flush=True)
    prog = bench.takara_program()
    for label, outs in (("no outputs", {}),
                        ("plain outputs", {"trimmed": [os.path.join(tmp, "o1.fq"), os.path.join(tmp, "o2.fq")]})):
        for rep in range(2):
            t0 = time.time()
            counters, timing = native.run_files(prog, ins, outs, gpus=1, threads=16)
            wall = time.time() - t0
        print(f"csq_run_files, {label}: {pairs / wall / 1e6:.2f} M pairs/s wall; read {timing.read_inflate:.2f} s, gpu {timing.h2d_kernels_d2h:.2f} s, "
              f"write {timing.write_deflate:.2f} s, total {timing.total:.2f} s", flush=True)


if __name__ == "__main__":
    main()
