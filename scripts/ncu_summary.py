#!/usr/bin/env python
"""Compact per-launch summary of an .ncu-rep (read on the CPU box): `python scripts/ncu_summary.py rep.ncu-rep [out.csv]`."""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "us"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "lsu_wb%"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ldg_req"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ldg_sectors"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "lsu_wavefronts"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": d["Kernel Name"].replace("<unnamed>::", "")[:70]}
        for key, short in COLS:
            v = d.get(key, "")
            u = units[hdr.index(key)] if key in hdr else ""
            try:
                f = float(v)
                if short.endswith("_MB") and u.lower().startswith("kbyte"):
                    f /= 1e3
                if short.endswith("_MB") and u.lower().startswith("gbyte"):
                    f *= 1e3
                if short == "us" and u == "ms":
                    f *= 1e3
                rec[short] = round(f, 2)
            except ValueError:
                rec[short] = v
        stalls = []
        for h in hdr:
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[h]), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        rec["top_stalls"] = " ".join(f"{n}={v:.1f}" for v, n in stalls[:4])
        out.append(rec)
    fields = ["kernel"] + [s for _, s in COLS] + ["top_stalls"]
    dst = open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout
    w = csv.DictWriter(dst, fieldnames=fields)
    w.writeheader()
    for rec in out:
        w.writerow(rec)


if __name__ == "__main__":
    main()
