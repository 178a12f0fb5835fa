#!/usr/bin/env python
"""Generate tests/golden/: inputs + expected outputs of the UNMODIFIED reference
``cutseq/run.py`` (imported from /root/reference) executed on the restated cutadapt in
oracle/cutadapt_shim.  Build-container only; the fixtures it writes are committed so the
GPU box (no /root/reference there) can check against them.

    python scripts/make_golden.py            # regenerate everything

PARITY UNPINNED: cutseq's own logic in these vectors is the reference's; the cutadapt
internals underneath are a restatement (see oracle/cutadapt_shim/cutadapt/__init__.py).
"""

import gzip
import hashlib
import io
import json
import os
import random
import shutil
import sys
import tempfile
from contextlib import redirect_stderr, redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
import run_reference  # noqa: E402

P5 = "ACACGACGCTCTTCCGATCT"
P7 = "AGATCGGAAGAGCACACGTC"
COMP = str.maketrans("ACGTN", "TGCAN")


def rc(s):
    return s.translate(COMP)[::-1]


def rand_seq(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def mutate(rng, s, sub=0.005, indel=0.0005, n_rate=0.001):
    out = []
    for ch in s:
        r = rng.random()
        if r < sub:
            out.append(rng.choice([c for c in "ACGT" if c != ch]))
        elif r < sub + indel:
            continue
        elif r < sub + 2 * indel:
            out.append(ch)
            out.append(rng.choice("ACGT"))
        elif r < sub + 2 * indel + n_rate:
            out.append("N")
        else:
            out.append(ch)
    return "".join(out)


def quals(rng, seq):
    q = [rng.choices("I9-", (0.84, 0.08, 0.08))[0] for _ in seq]
    if rng.random() < 0.2:
        tail = min(len(q), int(rng.expovariate(1 / 15.0)))
        for i in range(len(q) - tail, len(q)):
            q[i] = rng.choice("-#")
    for i, ch in enumerate(seq):
        if ch == "N":
            q[i] = "#"
    return "".join(q)


def write_fq(path, records):
    with gzip.open(path, "wb", compresslevel=9, mtime=0) if False else gzip.GzipFile(path, "wb", 9, mtime=0) as f:
        for name, seq, qual in records:
            f.write(f"@{name}\n{seq}\n+\n{qual}\n".encode())


def synth_pairs(seed, n, read_len=150, p5=P5, p7=P7, inline5="", inline3="", umi5=0, umi3=0, mask5=0, mask3=0,
                strand="-", readthrough=0.35, polya=0.05, suffix_style=None, bc_error=0.0, wrong_bc=0.0):
    """Fragments laid out as the library scheme says: R1 reads inline5+umi5+mask5+insert+mask3+umi3+inline3+P7...,
    R2 reads the reverse complement followed by rc(P5)."""
    rng = random.Random(seed)
    r1s, r2s = [], []
    for i in range(n):
        if rng.random() < readthrough:
            ins_len = int(rng.triangular(5, read_len - 1, read_len - 20))
        else:
            ins_len = rng.randint(read_len, 400)
        insert = rand_seq(rng, ins_len)
        if rng.random() < polya:
            run = rng.randint(10, 40)
            tail = "A" * run
            insert = (rc(tail) + insert) if strand == "-" else (insert + tail)
        b5, b3 = inline5, inline3
        if wrong_bc and rng.random() < wrong_bc:
            b5 = rand_seq(rng, len(b5))
            b3 = rand_seq(rng, len(b3))
        if bc_error:
            b5 = mutate(rng, b5, sub=bc_error, indel=bc_error / 8, n_rate=0)
            b3 = mutate(rng, b3, sub=bc_error, indel=bc_error / 8, n_rate=0)
        frag = b5 + rand_seq(rng, umi5) + rand_seq(rng, mask5) + insert + rand_seq(rng, mask3) + rand_seq(rng, umi3) + b3
        t1 = frag + p7 + "TGAACTCCAGTCAC" + "G" * read_len
        t2 = rc(frag) + rc(p5) + "AGATCTCGGTGGTCGCCGTATCATT" + "G" * read_len
        s1 = mutate(rng, t1[: read_len + 10])[:read_len]
        s2 = mutate(rng, t2[: read_len + 10])[:read_len]
        # occasional adapter dimers / 5' adapter remnants (template switching artefacts)
        if rng.random() < 0.01:
            s1 = (p5[rng.randint(0, 8):] + s1)[:read_len]
        if rng.random() < 0.01:
            s2 = (rc(p7)[rng.randint(0, 8):] + s2)[:read_len]
        if rng.random() < 0.003:
            s1 = s1[: rng.randint(0, 30)]
        tile, x, y = 1101 + i // 1000, rng.randint(1000, 30000), rng.randint(1000, 30000)
        base = f"SIM:1:FC:1:{tile}:{x}:{y}"
        if suffix_style == "slash":
            n1, n2 = base + "/1", base + "/2"
        elif suffix_style == "dot":
            n1, n2 = base + ".1", base + ".2"
        elif suffix_style == "bare":
            n1 = n2 = base
        elif suffix_style == "mixed":
            # mate numbers inside the id with a comment behind them: SuffixRemover (run.py:537-542) only strips at the
            # end of the whole header, so these ids keep their ".1" / "/1" and only match under dnaio's
            # record_names_match rule (one trailing 1-3 on both ids is not compared)
            style = i % 5
            if style == 0:
                n1, n2 = f"SRR1.{i + 1}.1 {i + 1} length={read_len}", f"SRR1.{i + 1}.2 {i + 1} length={read_len}"
            elif style == 1:
                n1, n2 = base + "/1 comment", base + "/2 comment"
            elif style == 2:
                n1, n2 = base + "/1", base + "/2"
            elif style == 3:
                n1, n2 = base + "\t1:N:0", base + "\t2:N:0"
            else:
                n1, n2 = base + "_1 x", base + "_2"
        else:
            n1, n2 = base + " 1:N:0:ACGTACGT+TGCATGCA", base + " 2:N:0:ACGTACGT+TGCATGCA"
        r1s.append((n1, s1, quals(rng, s1)))
        r2s.append((n2, s2, quals(rng, s2)))
    return r1s, r2s


def synth_smallrna(seed, n, read_len=75):
    rng = random.Random(seed)
    out = []
    ad = P7 + "ATCTCGTATGCCGTCTTCTGCTTG" + "G" * read_len
    for i in range(n):
        ins = rand_seq(rng, rng.randint(15, 45))
        tail = mutate(rng, ad[:read_len], sub=0.08, indel=0.01, n_rate=0.002)
        s = (mutate(rng, ins) + tail)[:read_len]
        if rng.random() < 0.02:
            s = ("AGTTCTACAGTCCGACGATC"[rng.randint(0, 9):] + s)[:read_len]
        name = f"SR:{i} extra comment" if i % 3 else f"SR:{i}/1"
        out.append((name, s, quals(rng, s)))
    return out


def edge_pairs():
    """Hand-made edge cases: empty reads, reads shorter than every cut, all-N, all low quality."""
    q = lambda s, c="I": c * len(s)
    recs = []

    def add(name, s1, s2, q1=None, q2=None):
        recs.append(((name + " 1:N:0:X", s1, q1 or q(s1)), (name + " 2:N:0:X", s2, q2 or q(s2))))

    add("e0", "", "")
    add("e1", "A", "C")
    add("e2", "ACGTACG", "TTTTTTTTTTTT")
    add("e3", "N" * 150, "N" * 150)
    add("e4", rand_seq(random.Random(1), 150), rand_seq(random.Random(2), 150), "#" * 150, "-" * 150)
    add("e5", P7, rc(P5))
    add("e6", P7[:3], rc(P5)[:3])
    add("e7", "ACGT" * 5 + P7[:2], "ACGT" * 5 + rc(P5)[:2])
    add("e8", P5 + "ACGTTGCA" * 6 + P7, rc(P7) + "ACGTTGCA" * 6 + rc(P5))
    add("e9", "T" * 150, "A" * 150)
    add("e10", "ACG" + "T" * 30 + rand_seq(random.Random(3), 60), rand_seq(random.Random(4), 60) + "A" * 30)
    add("e11", rand_seq(random.Random(5), 49), rand_seq(random.Random(6), 49))
    add("e12", rand_seq(random.Random(7), 50), rand_seq(random.Random(8), 50))
    add("e13", rand_seq(random.Random(9), 55), rand_seq(random.Random(10), 55))
    add("e14", (P7 + "A") * 7, (rc(P5) + "C") * 7)
    add("e15", "acgtacgtacgtacgtacgtacgt" + P7.lower(), "ACGTACGTACGTACGTACGTACGT" + rc(P5).lower())
    x = rand_seq(random.Random(11), 100)
    add("e16", x + P7[:10] + "T" + P7[10:], rc(x) + rc(P5)[:10] + rc(P5)[11:] + "ACGT")
    add("e17", x[:40] + P7[:19], rc(x[:40]) + rc(P5)[:19], "I" * 58 + "#", "I" * 30 + "#" * 29)
    return [r[0] for r in recs], [r[1] for r in recs]


def subset_bundled(n):
    out = []
    for mate in (1, 2):
        data = gzip.open(f"/root/reference/test/input_R{mate}.fq.gz").read().split(b"\n")
        recs = [(data[i][1:].decode(), data[i + 1].decode(), data[i + 3].decode()) for i in range(0, 4 * n, 4)]
        out.append(recs)
    return out


CASES = [
    # name, input set, argv
    ("takarav3", "bundled", ["-A", "TAKARAV3"]),
    ("takarav3_polya", "bundled", ["-A", "TAKARAV3", "--trim-polyA"]),
    ("takarav3_synth_polya", "synth_takara", ["-A", "TAKARAV3", "--trim-polyA"]),
    ("takarav3_polya_nodir_q15_m30", "synth_takara", ["-A", "TAKARAV3", "--trim-polyA", "--trim-polyA-wo-direction", "-q", "15", "-m", "30"]),
    ("takarav3_nocond", "synth_takara", ["-A", "TAKARAV3", "--no-conditional-cutter"]),
    ("takarav3_fmin80_anywhere", "synth_takara", ["-A", "TAKARAV3", "--force-trim-min-length", "80", "--force-anywhere"]),
    ("takarav3_autorc", "synth_takara", ["-A", "TAKARAV3", "--auto-rc"]),
    ("takarav3_edge", "edge", ["-A", "TAKARAV3", "--trim-polyA"]),
    ("sacseqv3_slash", "synth_sacseqv3", ["-A", "SACSEQV3", "--trim-polyA"]),
    ("inline_custom", "synth_inline", ["-a", "ACACGACGCTCTTCCGATCT(ATCACG)NNNNNNNNXXX<XXX(CGTGAT)AGATCGGAAGAGCACACGTC"]),
    ("inline_custom_ensure", "synth_inline", ["-a", "ACACGACGCTCTTCCGATCT(ATCACG)NNNNNNNNXXX<XXX(CGTGAT)AGATCGGAAGAGCACACGTC", "--ensure-inline-barcode"]),
    ("se_smallrna", "synth_se", ["-A", "SMALLRNA"]),
    ("se_inline_ensure", "synth_se_inline", ["-A", "INLINE", "--ensure-inline-barcode", "--trim-polyA"]),
    ("se_takarav3_autorc", "synth_takara_r1", ["-A", "TAKARAV3", "--auto-rc", "--trim-polyA"]),
    ("unstranded_polya", "synth_takara", ["-A", "UNSTRANDED", "--trim-polyA"]),
    ("takarav3_mate_numbers", "synth_names", ["-A", "TAKARAV3", "--trim-polyA"]),
]


def build_inputs():
    sets = {}
    sets["bundled"] = subset_bundled(1500)
    sets["synth_takara"] = synth_pairs(20240419, 700, mask5=3, mask3=6, umi3=8)
    sets["synth_takara_r1"] = [sets["synth_takara"][0]]
    sets["synth_sacseqv3"] = synth_pairs(7, 500, umi5=8, mask5=1, mask3=2, umi3=8, strand="+", suffix_style="slash")
    sets["synth_inline"] = synth_pairs(20240420, 700, inline5="ATCACG", inline3="CGTGAT", umi5=8, mask5=3, mask3=3,
                                       readthrough=0.3, bc_error=0.02, wrong_bc=0.03, suffix_style="dot")
    sets["synth_se"] = [synth_smallrna(20240421, 1500)]
    # INLINE scheme: AGTTCTACAGTCCGACGATCNNNNN>NNNNN(ATCACG)AGATCGGAAGAGCACACGTC, single-end
    r1, _ = synth_pairs(99, 600, read_len=100, p5="AGTTCTACAGTCCGACGATC", inline3="ATCACG", umi5=5, umi3=5, strand="+",
                        readthrough=0.7, bc_error=0.03, wrong_bc=0.05, suffix_style="bare")
    sets["synth_se_inline"] = [r1]
    sets["edge"] = edge_pairs()
    sets["synth_names"] = synth_pairs(31, 400, mask5=3, mask3=6, umi3=8, suffix_style="mixed")
    return sets


def main():
    if os.path.isdir(GOLD):
        shutil.rmtree(GOLD)
    os.makedirs(GOLD)
    sets = build_inputs()
    for name, mates in sets.items():
        for m, recs in enumerate(mates, 1):
            write_fq(os.path.join(GOLD, f"in_{name}_R{m}.fq.gz"), recs)
    manifest = []
    ref_run = run_reference.load_reference_main()
    for case, inset, argv in CASES:
        n_mates = len(sets[inset])
        with tempfile.TemporaryDirectory() as tmp:
            ins = [os.path.join(GOLD, f"in_{inset}_R{m}.fq.gz") for m in range(1, n_mates + 1)]
            sys.argv = ["cutseq"] + argv + ["-O", os.path.join(tmp, "out")] + ins
            err, outb = io.StringIO(), io.StringIO()
            with redirect_stderr(err), redirect_stdout(outb):
                ref_run.main()
            report = [l for l in err.getvalue().splitlines() if l.startswith(("status", "OK", "WARN"))]
            files = {}
            for fn in sorted(os.listdir(tmp)):
                data = gzip.open(os.path.join(tmp, fn)).read()
                key = fn[len("out_"):].replace(".fastq.gz", "")
                dst = os.path.join(GOLD, f"exp_{case}_{key}.fastq.gz")
                with gzip.GzipFile(dst, "wb", 9, mtime=0) as f:
                    f.write(data)
                files[key] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
            manifest.append({"case": case, "input": inset, "n_mates": n_mates, "argv": argv, "outputs": files, "minimal_report": report})
            print(case, {k: v["bytes"] for k, v in files.items()}, report[-1:] if report else "")
    # BASELINE.json config 1 = the reference's bundled 10 000 pairs: the two input DATA files are copied
    # byte for byte (the GPU box has no /root/reference); expected outputs are kept as hashes only
    for mate in (1, 2):
        shutil.copyfile(f"/root/reference/test/input_R{mate}.fq.gz", os.path.join(GOLD, f"in_bundled_full_R{mate}.fq.gz"))
    for case, argv in (("full_takarav3", ["-A", "TAKARAV3"]), ("full_takarav3_polya", ["-A", "TAKARAV3", "--trim-polyA"])):
        with tempfile.TemporaryDirectory() as tmp:
            sys.argv = ["cutseq"] + argv + ["-O", os.path.join(tmp, "out"), os.path.join(GOLD, "in_bundled_full_R1.fq.gz"), os.path.join(GOLD, "in_bundled_full_R2.fq.gz")]
            err = io.StringIO()
            with redirect_stderr(err):
                ref_run.main()
            files = {}
            for fn in sorted(os.listdir(tmp)):
                data = gzip.open(os.path.join(tmp, fn)).read()
                files[fn[len("out_"):].replace(".fastq.gz", "")] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
            report = [l for l in err.getvalue().splitlines() if l.startswith(("status", "OK", "WARN"))]
            manifest.append({"case": case, "input": "bundled_full", "hash_only": True, "n_mates": 2, "argv": argv, "outputs": files, "minimal_report": report})
            print(case, report[-1:])
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
