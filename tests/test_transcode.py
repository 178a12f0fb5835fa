"""``.bz2`` / ``.xz`` inputs and outputs (xopen's job in the reference): named pipes + Python codecs around the native
reader / writer.  No GPU: the native text reader reads the pipe, a stand-in for the library copies it to the outputs."""
import bz2
import lzma
import os
import threading

import pytest

from cutseq_b200 import native
from cutseq_b200.transcode import Transcoders

TEXT = b"".join(b"@r%d c\nACGTACGTACGTACGTNN\n+\nIIIIIIII99999999--\n" % i for i in range(20000))


def test_native_reader_takes_a_pipe(tmp_path):
    fifo = str(tmp_path / "p.fq")
    os.mkfifo(fifo)

    def feed():
        with open(fifo, "wb") as f:
            f.write(TEXT)

    t = threading.Thread(target=feed)
    t.start()
    got, total = [], 0
    with native.TextReader(fifo) as r:
        while True:
            n, texts, first = r.next(3000)
            if n == 0:
                break
            got.append(texts[0])
            total += n
    t.join()
    assert total == 20000 and b"".join(got) == TEXT


@pytest.mark.parametrize("ext,mod", [(".bz2", bz2), (".xz", lzma)])
def test_round_trip_through_the_pipes(tmp_path, ext, mod):
    src = str(tmp_path / ("in.fq" + ext))
    with mod.open(src, "wb") as f:
        f.write(TEXT)
    plain_in = str(tmp_path / "plain.fq")
    open(plain_in, "wb").write(TEXT)
    outs = {"trimmed": [str(tmp_path / ("o1.fq" + ext)), str(tmp_path / "o2.fq")], "short": [None, None], "untrimmed": None}
    with Transcoders([src, plain_in], outs) as tc:
        assert tc.inputs[1] == plain_in and tc.inputs[0] != src and tc.outputs["trimmed"][1] == outs["trimmed"][1]
        assert not tc.outputs["trimmed"][0].endswith(ext)
        # stand-in for csq_run_files: the native reader on the input pipe, plain writes to the output pipe
        with native.TextReader(tc.inputs[0]) as r, open(tc.outputs["trimmed"][0], "wb") as w:
            while True:
                n, texts, first = r.next(4096)
                if n == 0:
                    break
                w.write(texts[0])
    with mod.open(outs["trimmed"][0], "rb") as f:
        assert f.read() == TEXT
    assert not os.path.exists(os.path.dirname(tc.inputs[0]))  # the pipes and their directory are gone


def test_pumps_are_released_when_the_library_never_opens_its_pipes(tmp_path):
    src = str(tmp_path / "in.fq.bz2")
    with bz2.open(src, "wb") as f:
        f.write(TEXT)
    outs = {"trimmed": [str(tmp_path / "o.fq.xz")]}
    tc = Transcoders([src], outs)
    with pytest.raises(RuntimeError):
        with tc:
            raise RuntimeError("the library failed before it opened anything")
    assert len(tc._threads) == 2 and not any(t.is_alive() for t, _, _ in tc._threads)  # no pump is left waiting


def test_zstd_is_refused_with_a_message(tmp_path):
    with pytest.raises(ValueError) as e:
        Transcoders([str(tmp_path / "a.fq.zst")], {})
    assert "zstd" in str(e.value)


# ---- FASTA input: records without qualities go through converting pumps ----
def test_fasta_converters():
    import io

    from cutseq_b200 import transcode

    src = io.BytesIO(b">r1 desc\nACGT\nAC\n\n>r2\n\n>r3\r\nGG\r\n")
    dst = io.BytesIO()
    transcode.fasta_to_fastq(src, dst)
    assert dst.getvalue() == b"@r1 desc\nACGTAC\n+\nIIIIII\n@r2\n\n+\n\n@r3\nGG\n+\nII\n"
    back = io.BytesIO()
    transcode.fastq_to_fasta(io.BytesIO(dst.getvalue()), back)
    assert back.getvalue() == b">r1 desc\nACGTAC\n>r2\n\n>r3\nGG\n"
    with pytest.raises(ValueError):
        transcode.fasta_to_fastq(io.BytesIO(b"ACGT\n>r1\nAC\n"), io.BytesIO())


@pytest.mark.parametrize("ext", ["", ".gz", ".bz2"])
def test_fasta_files_go_in_and_come_out_as_fasta(tmp_path, ext):
    import gzip

    from cutseq_b200 import transcode

    fasta = b"".join(b">read%d some comment\nACGTACGTAC\nGTACGTNN\n" % i for i in range(5000))
    want = b"".join(b">read%d some comment\nACGTACGTACGTACGTNN\n" % i for i in range(5000))
    opener = {"": open, ".gz": gzip.open, ".bz2": bz2.open}[ext]
    src = str(tmp_path / ("in.fa" + ext))
    with opener(src, "wb") as f:
        f.write(fasta)
    assert transcode.is_fasta(src) and not transcode.is_fasta(__file__)
    outs = {"trimmed": [str(tmp_path / ("o1.fa" + ext))], "short": [str(tmp_path / "s1.fa")], "untrimmed": None}
    with Transcoders([src], outs) as tc:
        assert tc.fasta and tc.inputs[0] != src and tc.outputs["trimmed"][0] != outs["trimmed"][0]
        # stand-in for csq_run_files: FASTQ records arrive on the input pipe ...
        with native.TextReader(tc.inputs[0]) as r, open(tc.outputs["trimmed"][0], "wb") as w, open(tc.outputs["short"][0], "wb"):
            total = 0
            while True:
                n, texts, first = r.next(1024)
                if n == 0:
                    break
                assert texts[0].count(b"\n+\nIIIIIIIIIIIIIIIIII\n") == n  # constant qualities, one per base
                w.write(texts[0])  # ... and FASTQ records leave on the output pipes
                total += n
        assert total == 5000
    with opener(outs["trimmed"][0], "rb") as f:
        assert f.read() == want
    assert open(outs["short"][0], "rb").read() == b""


def test_mixed_fasta_and_fastq_inputs_are_refused(tmp_path):
    a, b = tmp_path / "a.fa", tmp_path / "b.fq"
    a.write_bytes(b">r\nACGT\n")
    b.write_bytes(b"@r\nACGT\n+\nIIII\n")
    with pytest.raises(ValueError) as e:
        Transcoders([str(a), str(b)], {})
    assert "differ in format" in str(e.value)


def test_zstd_goes_through_the_program_when_there_is_one(tmp_path, monkeypatch):
    """xopen falls back to the `zstd` program; the plumbing is tested with a stand-in program of that name (gzip behind
    zstd's command line: -dc file / -c)."""
    import gzip
    import stat

    fake = tmp_path / "bin" / "zstd"
    fake.parent.mkdir()
    fake.write_text("#!/bin/sh\nif [ \"$2\" = \"-dc\" ]; then exec gzip -dc \"$3\"; else exec gzip -1 -c; fi\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(fake.parent) + os.pathsep + os.environ["PATH"])
    src = str(tmp_path / "in.fq.zst")
    with gzip.open(src, "wb") as f:
        f.write(TEXT)
    outs = {"trimmed": [str(tmp_path / "o.fq.zst")], "short": [None], "untrimmed": None}
    with Transcoders([src], outs) as tc:
        with native.TextReader(tc.inputs[0]) as r, open(tc.outputs["trimmed"][0], "wb") as w:
            while True:
                n, texts, first = r.next(4096)
                if n == 0:
                    break
                w.write(texts[0])
    assert gzip.open(outs["trimmed"][0]).read() == TEXT
