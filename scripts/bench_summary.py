#!/usr/bin/env python
"""One line per bench JSON: value, step time, per-kernel sums, e2e.  usage: bench_summary.py file..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e)
        continue
    print(f, "value %.1fM ms %.3f" % (d["value"] / 1e6, d["ms_per_step"]), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["config"].get("numa_node"))
    ks = {}
    for k in d["kernels"]:
        ks[k["kernel"]] = ks.get(k["kernel"], 0) + k["ms"]
    print("   ", {k: round(v, 3) for k, v in ks.items()}, "sum %.2f" % sum(ks.values()))
    if d.get("e2e"):
        print("    e2e %.1fM %.2f ms" % (d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), round(d["roofline"]["achieved"]))
    if d.get("files"):
        print("    files", {k: (round(v.get("pairs_per_s")) if isinstance(v, dict) else v) for k, v in d["files"].items() if k in ("plain", "gz")}, "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]))
