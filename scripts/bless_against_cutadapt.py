#!/usr/bin/env python
"""Parity-pinning kit: run the REAL reference (cutseq's run.py on a REAL cutadapt 5.x + dnaio + xopen) over every
input under tests/golden/ and compare, byte for byte after decompression, with the committed expectations - which
were produced by the in-repo restatement (oracle/cutadapt_shim, scripts/make_golden.py) because no cutadapt can be
installed in the build image.

    python scripts/bless_against_cutadapt.py            # exit 0: every case identical; 1: differences; 3: no cutadapt
    python scripts/bless_against_cutadapt.py --write    # additionally record the outcome in tests/golden/PINNING.json

Where the reference comes from, first hit wins: $CUTSEQ_REFERENCE, /root/reference, baseline/_ref (a pip --target
install of the reference), an importable `cutseq` package.  The shim directory is never put on sys.path here, and a
`cutadapt` that resolves into oracle/ is rejected.
"""

from __future__ import annotations

import argparse
import gzip
import hashlib
import importlib
import importlib.metadata
import io
import json
import os
import sys
import tempfile
from contextlib import redirect_stderr, redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PINNING = os.path.join(GOLD, "PINNING.json")


def real_cutadapt():
    """-> (module, version) of an importable cutadapt that is not the in-repo shim, or (None, reason)."""
    shim = os.path.realpath(os.path.join(ROOT, "oracle"))
    sys.path[:] = [p for p in sys.path if not os.path.realpath(p or ".").startswith(shim)]
    for name in [m for m in sys.modules if m == "cutadapt" or m.startswith("cutadapt.")]:
        del sys.modules[name]
    try:
        mod = importlib.import_module("cutadapt")
    except Exception as exc:  # ImportError, or a broken install
        return None, f"import cutadapt failed: {exc!r}"
    where = os.path.realpath(getattr(mod, "__file__", "") or "")
    if where.startswith(shim):
        return None, f"cutadapt resolves to the in-repo shim ({where})"
    try:
        import dnaio  # noqa: F401
        import xopen  # noqa: F401
    except Exception as exc:
        return None, f"cutadapt found but dnaio/xopen missing: {exc!r}"
    return mod, getattr(mod, "__version__", "?")


def reference_main():
    for cand in (os.environ.get("CUTSEQ_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.exists(os.path.join(cand, "cutseq", "run.py")):
            sys.path.insert(0, cand)
            break
    real_version = importlib.metadata.version

    def version(name):  # run.py:190 asks for the installed version of "cutseq"
        try:
            return real_version(name)
        except importlib.metadata.PackageNotFoundError:
            if name in ("cutseq", "cutseq.run"):
                return "0.0.68"
            raise

    importlib.metadata.version = version
    import cutseq.run as ref_run

    return ref_run


def run_case(ref_run, case):
    n_mates = case["n_mates"]
    ins = [os.path.join(GOLD, f"in_{case['input']}_R{m}.fq.gz") for m in range(1, n_mates + 1)]
    with tempfile.TemporaryDirectory() as tmp:
        sys.argv = ["cutseq"] + list(case["argv"]) + ["-O", os.path.join(tmp, "out")] + ins
        err, out = io.StringIO(), io.StringIO()
        with redirect_stderr(err), redirect_stdout(out):
            ref_run.main()
        got = {}
        for fn in sorted(os.listdir(tmp)):
            got[fn[len("out_"):].replace(".fastq.gz", "")] = gzip.open(os.path.join(tmp, fn)).read()
        report = [l for l in err.getvalue().splitlines() if l.startswith(("status", "OK", "WARN"))]
    return got, report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true", help="record the outcome in tests/golden/PINNING.json")
    args = ap.parse_args()
    mod, info = real_cutadapt()
    if mod is None:
        print(f"bless: no real cutadapt here ({info}); tests/golden stays UNPINNED", file=sys.stderr)
        return 3
    ref_run = reference_main()
    with open(os.path.join(GOLD, "manifest.json")) as f:
        manifest = json.load(f)
    failures = []
    for case in manifest:
        got, report = run_case(ref_run, case)
        for key, want in case["outputs"].items():
            data = got.get(key)
            if data is None:
                failures.append(f"{case['case']}: reference wrote no {key}")
            elif hashlib.sha256(data).hexdigest() != want["sha256"]:
                line = next((i for i, (a, b) in enumerate(zip(data.split(b"\n"), (gzip.open(os.path.join(
                    GOLD, f"exp_{case['case']}_{key}.fastq.gz")).read() if not case.get("hash_only") else b"").split(b"\n"))) if a != b), None)
                failures.append(f"{case['case']}/{key}: bytes differ ({len(data)} vs {want['bytes']}; first differing line: {line})")
        for key in got:
            if key not in case["outputs"]:
                failures.append(f"{case['case']}: reference wrote an extra file {key}")
        if report[-1:] != case["minimal_report"][-1:]:
            failures.append(f"{case['case']}: minimal_report differs: {report[-1:]} vs {case['minimal_report'][-1:]}")
        print(("ok   " if not any(x.startswith(case["case"]) for x in failures) else "DIFF ") + case["case"])
    versions = {"cutadapt": info}
    for pkg in ("dnaio", "xopen"):
        try:
            versions[pkg] = importlib.metadata.version(pkg)
        except Exception:
            versions[pkg] = "?"
    if args.write:
        with open(PINNING, "w") as f:
            json.dump({"pinned": not failures, "checked_against": versions, "failures": failures,
                       "generator": "scripts/make_golden.py on oracle/cutadapt_shim"}, f, indent=1)
            f.write("\n")
    for x in failures:
        print("bless:", x, file=sys.stderr)
    print(f"bless: {len(manifest)} cases against cutadapt {info}: " + ("ALL IDENTICAL" if not failures else f"{len(failures)} differences"))
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
