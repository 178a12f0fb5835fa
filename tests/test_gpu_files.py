"""End-to-end through the CLI surface (cutseq_b200.run.main -> csq_run_files): FASTQ(.gz) files in, files out,
compared with the files the unmodified reference wrote (tests/golden, decompressed content)."""

import gzip
import os

import pytest

from cutseq_b200 import run
from tests import helpers

pytestmark = pytest.mark.gpu


def read_maybe_gz(path):
    with open(path, "rb") as f:
        head = f.read(2)
    return gzip.open(path).read() if head == b"\x1f\x8b" else open(path, "rb").read()


@pytest.mark.parametrize("case", helpers.golden_cases(), ids=lambda c: c["case"])
def test_cli_reproduces_reference_files(case, tmp_path, capsys):
    prefix = str(tmp_path / "out")
    run.main(case["argv"] + ["-O", prefix, "--batch-reads", "257"] + helpers.golden_input_paths(case))
    err = capsys.readouterr().err
    for key, meta in case["outputs"].items():
        got = read_maybe_gz(f"{prefix}_{key}.fastq.gz")
        assert got == helpers.golden_expected(case, key), key
    produced = sorted(f for f in os.listdir(tmp_path))
    assert produced == sorted(f"out_{k}.fastq.gz" for k in case["outputs"])
    if case["minimal_report"]:
        assert case["minimal_report"][-1] in err.splitlines()


def test_plain_outputs_and_explicit_names(tmp_path):
    case = [c for c in helpers.golden_cases() if c["case"] == "takarav3_synth_polya"][0]
    o1, o2, s1, s2 = (str(tmp_path / n) for n in ("t1.fq", "t2.fq", "s1.fq.gz", "s2.fq"))
    run.main(case["argv"] + ["-o", o1, o2, "-s", s1, s2, "-t", "3"] + helpers.golden_input_paths(case))
    assert open(o1, "rb").read() == helpers.golden_expected(case, "trimmed_R1")
    assert open(o2, "rb").read() == helpers.golden_expected(case, "trimmed_R2")
    assert gzip.open(s1).read() == helpers.golden_expected(case, "short_R1")
    assert open(s2, "rb").read() == helpers.golden_expected(case, "short_R2")


def test_bz2_and_xz_files_through_the_cli(tmp_path):
    """xopen's other formats: .bz2 input, .xz / .bz2 outputs (named pipes + Python codecs around the native reader /
    writer, cutseq_b200/transcode.py); a .zst path is refused."""
    import bz2
    import lzma

    case = [c for c in helpers.golden_cases() if c["case"] == "takarav3_synth_polya"][0]
    ins = []
    for m, p in enumerate(helpers.golden_input_paths(case)):
        q = str(tmp_path / f"in_R{m + 1}.fq.bz2")
        with bz2.open(q, "wb") as f:
            f.write(gzip.open(p).read())
        ins.append(q)
    o1, o2, s1, s2 = (str(tmp_path / n) for n in ("t1.fq.xz", "t2.fq.bz2", "s1.fq.gz", "s2.fq.xz"))
    run.main(case["argv"] + ["-o", o1, o2, "-s", s1, s2, "--batch-reads", "300"] + ins)
    assert lzma.open(o1).read() == helpers.golden_expected(case, "trimmed_R1")
    assert bz2.open(o2).read() == helpers.golden_expected(case, "trimmed_R2")
    assert gzip.open(s1).read() == helpers.golden_expected(case, "short_R1")
    assert lzma.open(s2).read() == helpers.golden_expected(case, "short_R2")
    with pytest.raises(ValueError):
        run.main(case["argv"] + ["-O", str(tmp_path / "z")] + [str(tmp_path / "a.fq.zst"), str(tmp_path / "b.fq.zst")])


def test_fasta_files_through_the_cli(tmp_path):
    """Records without qualities (reference run.py:439, 756 hand has_qualities() on to the writers): FASTA in, FASTA out,
    the same reads as a FASTQ run whose qualities are far above the cutoff."""
    import io

    from cutseq_b200 import transcode

    case = [c for c in helpers.golden_cases() if c["case"] == "takarav3_synth_polya"][0]
    fq, fa = [], []
    for m, p in enumerate(helpers.golden_input_paths(case)):
        lines = gzip.open(p).read().split(b"\n")
        recs = [(lines[i][1:], lines[i + 1]) for i in range(0, len(lines) - 1, 4)]
        q = str(tmp_path / f"in_R{m + 1}.fq")
        with open(q, "wb") as f:
            f.write(b"".join(b"@" + n + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for n, s in recs))
        a = str(tmp_path / f"in_R{m + 1}.fa.gz")
        with gzip.open(a, "wb") as f:  # sequences wrapped at 60 columns, as FASTA files often are
            f.write(b"".join(b">" + n + b"\n" + b"\n".join(s[o:o + 60] for o in range(0, max(len(s), 1), 60)) + b"\n" for n, s in recs))
        fq.append(q)
        fa.append(a)
    want = [str(tmp_path / n) for n in ("w1.fq", "w2.fq", "ws1.fq", "ws2.fq")]
    run.main(case["argv"] + ["-o", want[0], want[1], "-s", want[2], want[3], "--batch-reads", "500"] + fq)
    got = [str(tmp_path / n) for n in ("g1.fa", "g2.fa.gz", "gs1.fa", "gs2.fa.bz2")]
    run.main(case["argv"] + ["-o", got[0], got[1], "-s", got[2], got[3], "--batch-reads", "500"] + fa)
    import bz2

    for w, g, opener in zip(want, got, (open, gzip.open, open, bz2.open)):
        expect = io.BytesIO()
        transcode.fastq_to_fasta(io.BytesIO(open(w, "rb").read()), expect)
        assert opener(g, "rb").read() == expect.getvalue(), g
    assert open(want[0], "rb").read().count(b"\n") > 2000  # (the run did trim and write reads: 695 of the 700 pairs)


def test_batch_size_does_not_change_output(tmp_path):
    case = [c for c in helpers.golden_cases() if c["case"] == "inline_custom_ensure"][0]
    outs = []
    for i, br in enumerate(("1", "64", "100000")):
        prefix = str(tmp_path / f"o{i}")
        run.main(case["argv"] + ["-O", prefix, "--batch-reads", br] + helpers.golden_input_paths(case))
        outs.append({k: read_maybe_gz(f"{prefix}_{k}.fastq.gz") for k in case["outputs"]})
    assert outs[0] == outs[1] == outs[2]
    assert outs[0]["untrimmed_R1"] == helpers.golden_expected(case, "untrimmed_R1")


def test_json_report(tmp_path):
    import json

    case = helpers.golden_cases()[0]
    jf = str(tmp_path / "r.json")
    run.main(case["argv"] + ["-O", str(tmp_path / "o"), "--json-file", jf] + helpers.golden_input_paths(case))
    d = json.load(open(jf))
    fields = case["minimal_report"][-1].split("\t")
    assert d["read_counts"]["input"] == int(fields[1]) and d["read_counts"]["output"] == int(fields[6])
    assert d["barcode"]["p5"] == "ACACGACGCTCTTCCGATCT" and d["input"]["paired"] is True


def test_malformed_input_fails_loudly(tmp_path):
    from cutseq_b200 import native

    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"@r1\nACGT\n+\nII\n")
    with pytest.raises(native.NativeError):
        run.main(["-A", "SMALLRNA", "-O", str(tmp_path / "o"), str(bad)])


@pytest.mark.parametrize("case", helpers.golden_cases(hash_only=True), ids=lambda c: c["case"])
def test_config1_full_bundled_input(case, tmp_path, capsys):
    """BASELINE.json config 1: the reference's bundled 10 000 pairs through the CLI; outputs by sha256."""
    import hashlib

    prefix = str(tmp_path / "out")
    run.main(case["argv"] + ["-O", prefix, "-t", "4"] + helpers.golden_input_paths(case))
    err = capsys.readouterr().err
    for key, meta in case["outputs"].items():
        got = read_maybe_gz(f"{prefix}_{key}.fastq.gz")
        assert len(got) == meta["bytes"] and hashlib.sha256(got).hexdigest() == meta["sha256"], key
    assert case["minimal_report"][-1] in err.splitlines()


def test_multi_gpu_file_run_is_order_preserving(tmp_path):
    """csq_run_files sharding batches over 2 GPUs must give the single-GPU bytes (skipped on a 1-GPU box)."""
    from cutseq_b200 import native

    if native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    case = [c for c in helpers.golden_cases() if c["case"] == "takarav3_polya"][0]
    prefix = str(tmp_path / "out")
    run.main(case["argv"] + ["-O", prefix, "--gpus", "2", "--batch-reads", "97", "-t", "2"] + helpers.golden_input_paths(case))
    for key in case["outputs"]:
        assert read_maybe_gz(f"{prefix}_{key}.fastq.gz") == helpers.golden_expected(case, key), key


def _as_bgzf(src_gz, dst, piece=0xFF00, strip_final_newline=False, eof_marker=True):
    from scripts.bench_files import BGZF_EOF, bgzf_member

    text = gzip.open(src_gz).read()
    if strip_final_newline:
        text = text.rstrip(b"\n")
    with open(dst, "wb") as f:
        for o in range(0, len(text), piece):
            f.write(bgzf_member(text[o : o + piece], 6))
        if eof_marker:
            f.write(BGZF_EOF)
    return dst


@pytest.mark.parametrize("batch_reads,piece", [("257", 0xFF00), ("100000", 0xFF00), ("64", 3000), ("1000", 70)])
def test_bgzf_inputs_take_the_device_inflate_path(tmp_path, batch_reads, piece, capsys):
    """bgzip-style inputs: member runs go to the device (csq_submit_bgzf), batches are cut from the line counts the
    GPU returns (csq_bgzf_count_lines); pieces of 70 bytes put several members inside every record."""
    case = [c for c in helpers.golden_cases() if c["case"] == "takarav3_polya"][0]
    ins = [_as_bgzf(p, str(tmp_path / f"in_R{m + 1}.fq.gz"), piece) for m, p in enumerate(helpers.golden_input_paths(case))]
    prefix = str(tmp_path / "out")
    run.main(case["argv"] + ["-O", prefix, "--batch-reads", batch_reads] + ins)
    err = capsys.readouterr().err
    for key in case["outputs"]:
        assert read_maybe_gz(f"{prefix}_{key}.fastq.gz") == helpers.golden_expected(case, key), key
    assert case["minimal_report"][-1] in err.splitlines()


def test_bgzf_and_plain_inputs_without_a_final_line_end(tmp_path):
    case = [c for c in helpers.golden_cases() if c["case"] == "takarav3_synth_polya"][0]
    ins = [_as_bgzf(p, str(tmp_path / f"in_R{m + 1}.fq.gz"), strip_final_newline=True, eof_marker=(m == 0))
           for m, p in enumerate(helpers.golden_input_paths(case))]
    prefix = str(tmp_path / "out")
    run.main(case["argv"] + ["-O", prefix, "--batch-reads", "199"] + ins)
    for key in case["outputs"]:
        assert read_maybe_gz(f"{prefix}_{key}.fastq.gz") == helpers.golden_expected(case, key), key
    plain = []
    for m, p in enumerate(helpers.golden_input_paths(case)):
        q = str(tmp_path / f"in_R{m + 1}.fq")
        with open(q, "wb") as f:
            f.write(gzip.open(p).read().rstrip(b"\n") if m == 1 else gzip.open(p).read() + b"\n\n")
        plain.append(q)
    o1, o2 = str(tmp_path / "p1.fq"), str(tmp_path / "p2.fq")
    run.main(case["argv"] + ["-o", o1, o2, "--batch-reads", "199", "-t", "5"] + plain)
    assert open(o1, "rb").read() == helpers.golden_expected(case, "trimmed_R1")
    assert open(o2, "rb").read() == helpers.golden_expected(case, "trimmed_R2")


def test_mismatched_pair_files_and_corrupt_members_fail_and_leave_no_outputs(tmp_path):
    from cutseq_b200 import native

    case = [c for c in helpers.golden_cases() if c["case"] == "takarav3_synth_polya"][0]
    p1, p2 = helpers.golden_input_paths(case)
    a = _as_bgzf(p1, str(tmp_path / "a.fq.gz"))
    text2 = gzip.open(p2).read()
    short = str(tmp_path / "b.fq.gz")
    from scripts.bench_files import bgzf_member

    with open(short, "wb") as f:
        f.write(bgzf_member(b"\n".join(text2.split(b"\n")[:400]) + b"\n", 6))
    with pytest.raises(native.NativeError):
        run.main(case["argv"] + ["-O", str(tmp_path / "o"), a, short])
    assert not [f for f in os.listdir(tmp_path) if f.startswith("o_")]
    bad = bytearray(open(a, "rb").read())
    bad[len(bad) // 2] ^= 0xFF
    bad[len(bad) // 2 + 7] ^= 0xFF
    (tmp_path / "bad.fq.gz").write_bytes(bytes(bad))
    with pytest.raises(native.NativeError):
        run.main(case["argv"] + ["-O", str(tmp_path / "o2"), str(tmp_path / "bad.fq.gz"), _as_bgzf(p2, str(tmp_path / "b2.fq.gz"))])
    assert not [f for f in os.listdir(tmp_path) if f.startswith("o2_")]


@pytest.mark.parametrize("kind", ["plain", "bgzf"])
def test_multi_gpu_runs_of_the_fast_paths_are_order_preserving(tmp_path, kind):
    from cutseq_b200 import native

    if native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    case = [c for c in helpers.golden_cases() if c["case"] == "takarav3_polya"][0]
    ins = []
    for m, p in enumerate(helpers.golden_input_paths(case)):
        if kind == "bgzf":
            ins.append(_as_bgzf(p, str(tmp_path / f"in_R{m + 1}.fq.gz"), 4000))
        else:
            q = str(tmp_path / f"in_R{m + 1}.fq")
            open(q, "wb").write(gzip.open(p).read())
            ins.append(q)
    prefix = str(tmp_path / "out")
    run.main(case["argv"] + ["-O", prefix, "--gpus", "2", "--batch-reads", "61", "-t", "4"] + ins)
    for key in case["outputs"]:
        assert read_maybe_gz(f"{prefix}_{key}.fastq.gz") == helpers.golden_expected(case, key), key
