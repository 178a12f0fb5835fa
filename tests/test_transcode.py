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
