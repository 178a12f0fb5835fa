#!/usr/bin/env python
"""profiles/r02_traffic.json from an ncu summary of one whole chain (scripts/gpu_profile.sh -> chain_summary.csv):
DRAM bytes per pair of the emitter and executed integer lane-ops per read of the DP kernels, stamped with the hash of
the kernel sources that were profiled (bench.py only prints these numbers next to live times of the same sources).

    python scripts/make_traffic.py gpurun_out/<tag>/chain_summary.csv gpurun_out/<tag>/sources_hash.txt [pairs]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    summary, hash_file = sys.argv[1], sys.argv[2]
    pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
    rows = list(csv.DictReader(open(summary)))
    src_hash = open(hash_file).read().strip()
    # one step = the launches from the first k_nl_count up to and including the first k_emit_stage
    start = next(i for i, r in enumerate(rows) if r["kernel"].startswith("k_nl_count"))
    stop = next(i for i, r in enumerate(rows) if "k_emit_stage" in r["kernel"])
    step = rows[start:stop + 1]
    where = f"profiles/{os.path.basename(summary)}" if os.path.dirname(os.path.abspath(summary)) == os.path.join(ROOT, "profiles") else summary
    emit = step[-1]
    out = {
        "sources_sha256_16": src_hash,
        "source": f"ncu --set full --clock-control none of one chain at {pairs} pairs ({where})",
        "k_emit": {"dram_bytes_per_pair": (float(emit["dram_rd_MB"]) + float(emit["dram_wr_MB"])) * 1e6 / pairs, "pairs_per_launch_measured": pairs},
        "chain_dram_bytes_per_pair": sum(float(r["dram_rd_MB"]) + float(r["dram_wr_MB"]) for r in step) * 1e6 / pairs,
        "chain_launches_per_step": len(step),
        "dp_lane_ops_per_read": {},
    }
    # the ALIGN ops in program order per mate: rightmost_front, back, then the homopolymer op of the mate
    names = {0: ["rightmost_front", "back", "noninternal_front"], 1: ["rightmost_front", "back", "noninternal_back"]}
    mate, seen = -1, 0
    for r in step:
        k = r["kernel"]
        if k.startswith("k_nl_count"):
            mate, seen = mate + 1, 0
        lane_ops = float(r["warp_inst"]) * 32 / pairs
        if "k_prefilter" in k and mate in names:
            out["dp_lane_ops_per_read"].setdefault(f"k_prefilter({names[mate][seen]})", lane_ops)
        elif "k_align" in k and mate in names:
            out["dp_lane_ops_per_read"].setdefault(f"k_align({names[mate][seen]})", lane_ops)
            seen += 1
    dst = os.path.join(ROOT, "profiles", "r02_traffic.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
