#!/usr/bin/env python
"""What the GPU box's host can do for the whole-file path: page cache -> pinned memory (pread) and pinned memory ->
page cache (pwrite) by thread count, pinned vs pageable destination, and the PCIe copy ceiling (csq_pcie_peak).

    python scripts/host_io_probe.py --gb 4 --out gpurun_out/host_io_probe.json
"""

import argparse
import json
import os
import shutil
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_threads(fn, n):
    ths = [threading.Thread(target=fn, args=(i,)) for i in range(n)]
    t0 = time.perf_counter()
    [t.start() for t in ths]
    [t.join() for t in ths]
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--tmp", default=None)
    args = ap.parse_args()
    import numpy as np
    import torch

    from cutseq_b200 import native

    res = {"cpus": os.cpu_count(), "sched_cpus": len(os.sched_getaffinity(0))}
    try:
        with open("/proc/meminfo") as f:
            mi = {l.split(":")[0]: l.split(":")[1].strip() for l in f}
        res["mem_total"] = mi.get("MemTotal")
        res["mem_available"] = mi.get("MemAvailable")
    except OSError:
        pass
    tmp = tempfile.mkdtemp(dir=args.tmp)
    st = os.statvfs(tmp)
    res["tmp"] = {"path": tmp, "free_gb": st.f_bavail * st.f_frsize / 1e9}
    try:
        with open("/proc/mounts") as f:
            mounts = [l.split() for l in f]
        best = max((m for m in mounts if tmp.startswith(m[1])), key=lambda m: len(m[1]))
        res["tmp"]["fs"] = best[2]
    except Exception:
        pass
    n = int(args.gb * (1 << 30))
    pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    pinned.fill_(65)
    pv = memoryview(pinned.numpy())
    pageable = np.full(n, 66, dtype=np.uint8)
    gv = memoryview(pageable)
    path = os.path.join(tmp, "probe.bin")
    try:
        for label, view in (("pinned", pv), ("pageable", gv)):
            for T in (1, 2, 4, 8, 16, 32):
                if T > 2 * (os.cpu_count() or 1):
                    continue
                per = n // T
                fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)

                def wr(i):
                    off, end = i * per, (i + 1) * per
                    while off < end:
                        off += os.pwrite(fd, view[off:min(end, off + (8 << 20))], off)

                dt = run_threads(wr, T)
                os.close(fd)
                res[f"pwrite_{label}_T{T}_GBps"] = per * T / dt / 1e9
                fd = os.open(path, os.O_RDONLY)

                def rd(i):
                    off, end = i * per, (i + 1) * per
                    while off < end:
                        off += os.preadv(fd, [view[off:min(end, off + (8 << 20))]], off)

                dt = run_threads(rd, T)
                os.close(fd)
                res[f"pread_{label}_T{T}_GBps"] = per * T / dt / 1e9
                os.remove(path)
        for mode, name in ((0, "h2d_alone"), (1, "d2h_alone"), (2, "both")):
            a, b = native.pcie_peak(0, 1 << 30, 1 << 30, reps=6, mode=mode)
            res[f"pcie_{name}_GBps"] = [a, b]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(res, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
