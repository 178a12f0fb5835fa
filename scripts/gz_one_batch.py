#!/usr/bin/env python
"""One BGZF batch -> gzip through the chain (for ncu launch lists of the gzip kernels): python scripts/gz_one_batch.py --pairs 1000000"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser(); ap.add_argument("--pairs", type=int, default=1_000_000); ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
import bench
from cutseq_b200 import _abi as A, native
from scripts import bench_files
prog = bench.takara_program(); n = args.pairs
batch = native.synth_batch(2, n, first_index=0, buffer=0)
texts = [bytes(native.format_fastq(batch, m)) for m in range(2)]
runs = [native.BgzfRun(bench_files.bgzf_compress(t, os.cpu_count() or 8)) for t in texts]
with native.Plan(prog, 0, A.PLAN_GZIP_OUT) as plan:
    for _ in range(args.reps):
        plan.run_bgzf(runs, n, capacity=len(texts[0]) + 64 * n)
print("ok")
