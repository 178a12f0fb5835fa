"""Host FASTQ parser (csq_parse_fastq_mem / csq_reader_*) - runs without a GPU."""

import ctypes as C
import gzip
import os
import random

import numpy as np
import pytest

from cutseq_b200 import _abi as A
from cutseq_b200 import native
from tests import helpers


def parse_mem(text: bytes, max_reads=1 << 20):
    L = native.lib()
    cap = len(text) * 2 + 1024
    seq = np.zeros(cap, np.uint8)
    qual = np.zeros(cap, np.uint8)
    name = np.zeros(cap, np.uint8)
    n_max = text.count(b"\n") // 4 + 2
    seq_off = np.zeros(n_max, np.uint32)
    seq_len = np.zeros(n_max, np.uint32)
    name_off = np.zeros(n_max + 1, np.uint32)
    n, sb, used = C.c_uint32(), C.c_uint64(), C.c_uint64()
    buf = np.frombuffer(text + b"\0", dtype=np.uint8)
    rc = L.csq_parse_fastq_mem(buf.ctypes.data, len(text), max_reads, seq.ctypes.data, qual.ctypes.data, cap, seq_off.ctypes.data,
                               seq_len.ctypes.data, name.ctypes.data, cap, name_off.ctypes.data, C.byref(n), C.byref(sb), C.byref(used))
    if rc:
        raise native.NativeError(rc, L.csq_last_error().decode())
    recs = []
    for i in range(n.value):
        o, l = int(seq_off[i]), int(seq_len[i])
        assert o % 16 == 0
        recs.append((name[name_off[i]:name_off[i + 1]].tobytes().decode(), seq[o:o + l].tobytes().decode(), qual[o:o + l].tobytes().decode()))
    return recs, used.value


def test_parse_basic_and_edge_cases():
    recs, used = parse_mem(b"@r1 c\nACGT\n+\nIIII\n@r2\n\n+\n\n@r3\nAC\n+r3\n#I")
    assert recs == [("r1 c", "ACGT", "IIII"), ("r2", "", ""), ("r3", "AC", "#I")]
    recs, _ = parse_mem(b"@r1\r\nACGT\r\n+\r\nIIII\r\n")
    assert recs == [("r1", "ACGT", "IIII")]
    recs, _ = parse_mem(b"")
    assert recs == []
    recs, _ = parse_mem(b"@a\nA\n+\nI\n\n\n")
    assert recs == [("a", "A", "I")]
    recs, used = parse_mem(b"@a\nA\n+\nI\n@b\nC\n+\nI\n", max_reads=1)
    assert recs == [("a", "A", "I")] and used == 9


@pytest.mark.parametrize("text,code", [
    (b"r1\nACGT\n+\nIIII\n", A.ERR_FORMAT),
    (b"@r1\nACGT\n-\nIIII\n", A.ERR_FORMAT),
    (b"@r1\nACGT\n+\nIII\n", A.ERR_FORMAT),
    (b"@r1\nACGT\n+\n", A.ERR_FORMAT),
    (b"@r1\n" + b"A" * 900 + b"\n+\n" + b"I" * 900 + b"\n", A.ERR_LIMIT),
    (b"@r1 c\nACGT\n+r1\nIIII\n", A.ERR_FORMAT),   # dnaio: the second description must be empty or equal to the first
    (b"@" + b"h" * 70000 + b"\nACGT\n+\nIIII\n", A.ERR_LIMIT),
])
def test_parse_errors(text, code):
    with pytest.raises(native.NativeError) as e:
        parse_mem(text)
    assert e.value.code == code


def test_reader_matches_python_parse():
    L = native.lib()
    case = helpers.golden_cases()[0]
    paths = helpers.golden_input_paths(case)
    want = helpers.golden_inputs(case)
    h = C.c_void_p()
    native.check(L.csq_reader_open(os.fsencode(paths[0]), os.fsencode(paths[1]), C.byref(h)))
    got = [[], []]
    buf = 0
    while True:
        b = A.csq_batch_in()
        native.check(L.csq_reader_next(h, buf, 333, C.byref(b)))
        if b.n_reads == 0:
            break
        assert b.n_reads <= 333 and b.n_mates == 2
        for m in range(2):
            mi = b.mate[m]
            n = b.n_reads
            seq_off = np.ctypeslib.as_array(C.cast(mi.seq_off, C.POINTER(C.c_uint32)), (n,))
            seq_len = np.ctypeslib.as_array(C.cast(mi.seq_len, C.POINTER(C.c_uint32)), (n,))
            name_off = np.ctypeslib.as_array(C.cast(mi.name_off, C.POINTER(C.c_uint32)), (n + 1,))
            for i in range(n):
                got[m].append((C.string_at(mi.name + int(name_off[i]), int(name_off[i + 1] - name_off[i])).decode(),
                               C.string_at(mi.seq + int(seq_off[i]), int(seq_len[i])).decode(),
                               C.string_at(mi.qual + int(seq_off[i]), int(seq_len[i])).decode()))
        buf ^= 1
    L.csq_reader_close(h)
    assert got[0] == want[0] and got[1] == want[1]


def test_reader_rejects_unequal_pairs(tmp_path):
    p1, p2 = tmp_path / "a.fq", tmp_path / "b.fq.gz"
    p1.write_bytes(b"@a\nA\n+\nI\n@b\nA\n+\nI\n")
    with gzip.open(p2, "wb") as f:
        f.write(b"@a\nA\n+\nI\n")
    L = native.lib()
    h = C.c_void_p()
    native.check(L.csq_reader_open(os.fsencode(str(p1)), os.fsencode(str(p2)), C.byref(h)))
    b = A.csq_batch_in()
    assert L.csq_reader_next(h, 0, 100, C.byref(b)) == A.ERR_FORMAT
    L.csq_reader_close(h)


def test_missing_file_is_an_io_error():
    L = native.lib()
    h = C.c_void_p()
    assert L.csq_reader_open(b"/nonexistent/x.fq", None, C.byref(h)) == A.ERR_IO


def test_synth_generator_is_counter_based():
    from oracle import oracle  # noqa: F401  (only to make sure both sides read the same batch layout)

    a = native.synth_batch(2, 1000, first_index=500, buffer=0)
    seq_a = C.string_at(a.mate[0].seq, a.mate[0].seq_bytes)
    name_a = C.string_at(a.mate[1].name, a.mate[1].name_bytes)
    b = native.synth_batch(2, 1500, first_index=0, buffer=1)
    seq_b = C.string_at(b.mate[0].seq, b.mate[0].seq_bytes)
    assert seq_b[500 * 160:] == seq_a  # record i depends only on (seed, index)
    assert name_a.count(b" 2:N:0:") == 1000
    c = native.synth_batch(4, 100, buffer=2)
    assert c.n_mates == 1 and c.mate[0].seq_bytes == 100 * 80


def test_abi_exports_every_declared_symbol():
    import re

    hdr = open(os.path.join(helpers.ROOT, "include", "cutseq_b200.h")).read()
    names = set(re.findall(r"\b(csq_[a-z_0-9]+)\s*\(", hdr))
    names -= {"csq_plan", "csq_reader"}
    L = native.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert L.csq_abi_version() == A.ABI_VERSION


def test_struct_sizes_match_header(tmp_path):
    """Compile a tiny C program against include/cutseq_b200.h and compare sizeof() with the ctypes mirror."""
    import subprocess

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "cutseq_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(csq_op),sizeof(csq_filters),sizeof(csq_mate_in),sizeof(csq_batch_in),sizeof(csq_text_out),"
                   "sizeof(csq_batch_out),sizeof(csq_match),sizeof(csq_counters),sizeof(csq_files),sizeof(csq_timing));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(helpers.ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(t) for t in (A.csq_op, A.csq_filters, A.csq_mate_in, A.csq_batch_in, A.csq_text_out, A.csq_batch_out,
                                  A.csq_match, A.csq_counters, A.csq_files, A.csq_timing)]
    assert got == want


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the compute entry points must fail (no CPU fallback)."""
    try:
        n = native.device_count()
    except native.NativeError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    prog = helpers.program_for(["-A", "TAKARAV3"], 2)
    with pytest.raises(native.NativeError) as e:
        native.Plan(prog)
    assert e.value.code == A.ERR_NO_DEVICE


def test_host_placement_calls_are_safe_without_a_gpu():
    """csq_bind_host_to_device needs the device's PCI address: without a device it reports that and changes nothing;
    csq_unbind_host is always callable."""
    import os

    before = os.sched_getaffinity(0)
    try:
        have_gpu = native.device_count() > 0
    except native.NativeError:  # no driver at all
        have_gpu = False
    if not have_gpu:
        with pytest.raises(native.NativeError):
            native.bind_host_to_device(0)
    else:
        assert native.bind_host_to_device(0) >= -1
    native.unbind_host()
    assert os.sched_getaffinity(0) == before


# ---- text batches: the reader only cuts the byte stream at record boundaries (text_reader.cpp) ----
def _records(n, read_len=30, crlf=False, seed=5):
    rng = random.Random(seed)
    eol = "\r\n" if crlf else "\n"
    out = []
    for i in range(n):
        L = rng.randint(0, read_len)
        seq = "".join(rng.choice("ACGTN") for _ in range(L))
        qual = "".join(rng.choice("#-9I") for _ in range(L))
        out.append(f"@r{i} c{i}{eol}{seq}{eol}+{eol}{qual}{eol}")
    return out


def test_newline_helpers():
    L = native.lib()
    rng = random.Random(1)
    for n in (0, 1, 63, 64, 65, 4096, 4097, 100001):
        buf = np.frombuffer(bytes(rng.choice(b"ACGT\n\n") for _ in range(n)), dtype=np.uint8) if n else np.zeros(1, np.uint8)
        raw = buf.tobytes()[:n]
        assert L.csq_count_newlines(buf.ctypes.data, n) == raw.count(b"\n")
        for k in (1, 2, raw.count(b"\n"), raw.count(b"\n") + 1):
            pos, want = -1, 2 ** 64 - 1
            for _ in range(k):
                pos = raw.find(b"\n", pos + 1)
                if pos < 0:
                    break
            else:
                want = pos + 1
            if k >= 1:
                assert L.csq_after_kth_newline(buf.ctypes.data, n, k) == want, (n, k)


@pytest.mark.parametrize("gz", [False, True], ids=["plain", "gzip"])
@pytest.mark.parametrize("tail", ["", "nofinal", "blank", "blank_crlf"])
def test_text_reader_cuts_whole_records(tmp_path, gz, tail):
    recs1, recs2 = _records(1000, seed=1), _records(1000, seed=2, crlf=(tail == "blank_crlf"))
    t1, t2 = "".join(recs1), "".join(recs2)
    if tail == "nofinal":
        t1 = t1[:-1]
    elif tail == "blank":
        t1 += "\n\n"
    elif tail == "blank_crlf":
        t2 += "\r\n"
    p1, p2 = tmp_path / ("a.fq" + (".gz" if gz else "")), tmp_path / ("b.fq" + (".gz" if gz else ""))
    if gz:  # two concatenated members: one valid gzip file
        half = len(t1) // 2
        p1.write_bytes(gzip.compress(t1[:half].encode()) + gzip.compress(t1[half:].encode()))
        p2.write_bytes(gzip.compress(t2.encode()))
    else:
        p1.write_bytes(t1.encode())
        p2.write_bytes(t2.encode())
    for batch in (1, 7, 256, 5000):
        got1, got2, total, first = b"", b"", 0, 0
        with native.TextReader(str(p1), str(p2)) as r:
            while True:
                n, texts, first_record = r.next(batch)
                if n == 0:
                    break
                assert first_record == total and n <= batch
                assert texts[0].count(b"\n") == 4 * n and texts[1].count(b"\n") == 4 * n
                got1 += texts[0]
                got2 += texts[1]
                total += n
        assert total == 1000
        assert got1 == "".join(recs1).encode() and got2 == "".join(recs2).encode()


@pytest.mark.parametrize("gz", [False, True], ids=["plain", "gzip"])
def test_text_reader_reads_pipes(tmp_path, gz):
    """Inputs that cannot be mapped (named pipes, process substitution, stdin): plain text streams through read(),
    gzip data through zlib (xopen reads such inputs for the reference, run.py:434, 751)."""
    import threading

    recs = [_records(800, seed=5), _records(800, seed=6)]
    fifos = [tmp_path / "a.fifo", tmp_path / "b.fifo"]
    for f in fifos:
        os.mkfifo(f)

    def feed(path, text):
        data = gzip.compress(text.encode()) if gz else text.encode()
        with open(path, "wb") as w:
            w.write(data)

    pumps = [threading.Thread(target=feed, args=(str(f), "".join(r)), daemon=True) for f, r in zip(fifos, recs)]
    for t in pumps:
        t.start()
    got, total = [b"", b""], 0
    with native.TextReader(str(fifos[0]), str(fifos[1])) as r:
        while True:
            n, texts, _ = r.next(300)
            if n == 0:
                break
            got[0] += texts[0]
            got[1] += texts[1]
            total += n
    for t in pumps:
        t.join(timeout=10)
    assert total == 800
    assert got[0] == "".join(recs[0]).encode() and got[1] == "".join(recs[1]).encode()


@pytest.mark.parametrize("threads", ["1", "4", "7"])
def test_text_reader_parallel_pread_of_plain_files(tmp_path, monkeypatch, threads):
    """Plain regular files: several threads pread() pieces of a batch; batches of very different sizes in one file
    (the size estimate comes from the previous batch), a last batch that ends at the end of the file."""
    monkeypatch.setenv("CSQ_READ_THREADS", threads)
    recs = _records(30000, read_len=20, seed=3) + _records(30000, read_len=300, seed=4) + _records(5000, read_len=5, seed=6)
    text = "".join(recs).encode()
    assert len(text) > 10 << 20
    (tmp_path / "big.fq").write_bytes(text[:-1])  # no final line end
    for batch in (4096, 20000, 70000):
        got, total = [], 0
        with native.TextReader(str(tmp_path / "big.fq")) as r:
            while True:
                n, texts, first_record = r.next(batch)
                if n == 0:
                    break
                assert first_record == total and n <= batch and texts[0].count(b"\n") == 4 * n
                got.append(texts[0])
                total += n
        assert total == len(recs) and b"".join(got) == text


def _bgzf(data: bytes, block: int = 0xFF00) -> bytes:
    """bgzip's container: gzip members of at most 64 KiB with a 'BC' extra field (total size - 1), an empty member
    at the end (SAM/BAM specification, section 4.1)."""
    import struct
    import zlib

    out = b""
    for off in list(range(0, len(data), block)) + [None]:
        chunk = b"" if off is None else data[off:off + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(chunk) + c.flush()
        total = 12 + 6 + len(body) + 8
        out += (b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, total - 1)
                + body + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return out


@pytest.mark.parametrize("threads", ["1", "5"])
def test_text_reader_inflates_bgzf_members_side_by_side(tmp_path, monkeypatch, threads):
    monkeypatch.setenv("CSQ_READ_THREADS", threads)
    recs = _records(20000, read_len=150, seed=8) + _records(3000, read_len=5, seed=9)
    text = "".join(recs).encode()
    (tmp_path / "a.fq.gz").write_bytes(_bgzf(text))
    assert gzip.decompress((tmp_path / "a.fq.gz").read_bytes()) == text  # the fixture is a valid gzip file
    for batch in (1000, 9000, 50000):
        got, total = [], 0
        with native.TextReader(str(tmp_path / "a.fq.gz")) as r:
            while True:
                n, texts, first_record = r.next(batch)
                if n == 0:
                    break
                assert first_record == total and texts[0].count(b"\n") == 4 * n
                got.append(texts[0])
                total += n
        assert total == len(recs) and b"".join(got) == text
    # a flipped byte in the middle of the file is caught by the member's CRC / structure
    raw = bytearray(_bgzf(text))
    raw[len(raw) // 2] ^= 0x55
    (tmp_path / "bad.fq.gz").write_bytes(bytes(raw))
    with native.TextReader(str(tmp_path / "bad.fq.gz")) as r:
        with pytest.raises(native.NativeError) as e:
            while r.next(50000)[0]:
                pass
        assert e.value.code == A.ERR_IO


def test_text_reader_errors(tmp_path):
    (tmp_path / "short.fq").write_bytes(b"@r1\nACGT\n+\nIIII\n@r2\nAC\n")
    with native.TextReader(str(tmp_path / "short.fq")) as r:
        with pytest.raises(native.NativeError) as e:
            r.next(100)
        assert e.value.code == A.ERR_FORMAT and "prematurely" in str(e.value)
    (tmp_path / "a.fq").write_bytes("".join(_records(10)).encode())
    (tmp_path / "b.fq").write_bytes("".join(_records(9)).encode())
    with native.TextReader(str(tmp_path / "a.fq"), str(tmp_path / "b.fq")) as r:
        with pytest.raises(native.NativeError) as e:
            r.next(100)
        assert e.value.code == A.ERR_FORMAT
    (tmp_path / "trunc.fq.gz").write_bytes(gzip.compress("".join(_records(500)).encode())[:-20])
    with native.TextReader(str(tmp_path / "trunc.fq.gz")) as r:
        with pytest.raises(native.NativeError) as e:
            r.next(1000)
        assert e.value.code == A.ERR_IO
    (tmp_path / "empty.fq").write_bytes(b"")
    with native.TextReader(str(tmp_path / "empty.fq")) as r:
        assert r.next(10)[0] == 0


def test_format_fastq_round_trips_through_the_parser():
    batch = native.synth_batch(3, 500, first_index=7, buffer=4)
    for m in range(2):
        text = native.format_fastq(batch, m)
        assert text.tobytes().count(b"\n") == 4 * 500
        rec = parse_mem(text.tobytes())
        mi = batch.mate[m]
        noff = np.ctypeslib.as_array(C.cast(mi.name_off, C.POINTER(C.c_uint32)), (501,))
        assert rec[0][0][0] == C.string_at(mi.name, int(noff[1])).decode()
