#!/usr/bin/env python
"""Where does the time of the device gzip path go?  Wall-clock of the synchronous calls on one batch:
inflate + line count alone, BGZF batch -> text, text batch -> gzip, BGZF batch -> gzip.

    python scripts/gz_probe.py --pairs 500000
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=500_000)
    ap.add_argument("--level", type=int, default=1)
    args = ap.parse_args()
    import bench
    from cutseq_b200 import _abi as A
    from cutseq_b200 import native
    from scripts import bench_files

    prog = bench.takara_program()
    n = args.pairs
    batch = native.synth_batch(2, n, first_index=0, buffer=0)
    texts = [bytes(native.format_fastq(batch, m)) for m in range(2)]
    t0 = time.time()
    zs = [bench_files.bgzf_compress(t, os.cpu_count() or 8) for t in texts]
    print(f"text {sum(map(len, texts)) / 1e6:.1f} MB -> bgzf {sum(map(len, zs)) / 1e6:.1f} MB in {time.time() - t0:.2f} s (host zlib level 1)")
    runs = [native.BgzfRun(z) for z in zs]
    cap = len(texts[0]) + 64 * n

    def timed(label, fn, reps=3):
        best = 1e9
        for _ in range(reps):
            t0 = time.time()
            out = fn()
            best = min(best, time.time() - t0)
        print(f"{label:45s} {best * 1e3:9.2f} ms   {sum(map(len, texts)) / best / 1e9:7.2f} GB/s of text   {n / best / 1e6:7.2f} M pairs/s")
        return out

    with native.Plan(prog, 0, 0) as plan:
        timed("count_lines (H2D + inflate + CRC), one mate", lambda: plan.bgzf_count_lines(runs[0]))
        timed("text batch -> text", lambda: plan.run_text(texts, n, capacity=cap))
        timed("BGZF batch -> text", lambda: plan.run_bgzf(runs, n, capacity=cap))
    with native.Plan(prog, 0, A.PLAN_GZIP_OUT) as plan:
        z = timed("text batch -> gzip", lambda: plan.run_text(texts, n, capacity=cap))
        print("   gzip out bytes", sum(len(z[0][d][m]) for d in range(3) for m in range(2)), "of text", sum(map(len, texts)))
        timed("BGZF batch -> gzip", lambda: plan.run_bgzf(runs, n, capacity=cap))


if __name__ == "__main__":
    main()
