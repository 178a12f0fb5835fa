"""The library's own gzip / DEFLATE decoder (csrc/inflate.cpp, csq_gunzip_mem) against zlib - runs without a GPU.

Every block type (stored, fixed, dynamic), every resumption point of the piece-wise API (pieces of 1 byte to
1 MiB, i.e. output buffers that end inside a match, right at a block end, ...), concatenated members, header
flags, small windows, truncation and single-bit corruption."""

import ctypes as C
import gzip
import io
import os
import random
import zlib

import numpy as np
import pytest

from cutseq_b200 import native


def gunzip(data: bytes, piece: int, cap: int):
    L = native.lib()
    src = np.frombuffer(data + b"\0" * 16, dtype=np.uint8)
    dst = np.zeros(max(cap, 1), dtype=np.uint8)
    n = C.c_uint64()
    rc = L.csq_gunzip_mem(src.ctypes.data, len(data), dst.ctypes.data, cap, piece, C.byref(n))
    if rc:
        raise native.NativeError(rc, L.csq_last_error().decode())
    return dst[: n.value].tobytes()


def gz(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=31):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    return c.compress(data) + c.flush()


def fastq(n, seed=3):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        ln = rng.randint(20, 150)
        out.append(f"@read{i} {rng.randint(0, 99999)}\n{''.join(rng.choice('ACGT') for _ in range(ln))}\n+\n"
                   f"{''.join(rng.choice('I9-#') for _ in range(ln))}\n")
    return "".join(out).encode()


CASES = {
    "empty": b"", "one": b"a", "text": b"hello world\n" * 1000, "random": os.urandom(60000), "fastq": fastq(2000),
    "zeros": bytes(70000), "period2": b"ab" * 40000, "period3": b"abc" * 30000, "period7": b"abcdefg" * 15000,
    "bytes": bytes(range(256)) * 200,
}


@pytest.mark.parametrize("name", list(CASES))
def test_round_trip_all_block_types_and_piece_sizes(name):
    data = CASES[name]
    for level in (0, 1, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
            comp = gz(data, level, strategy)
            for piece in (1, 7, 300, 319, 320, 321, 4096, 1 << 20):
                if piece < 300 and len(data) > 50000 and (level, strategy) != (6, zlib.Z_DEFAULT_STRATEGY):
                    continue
                assert gunzip(comp, piece, len(data) + 10) == data, (name, level, strategy, piece)


def test_members_header_flags_padding_and_windows():
    buf = io.BytesIO()
    with gzip.GzipFile(filename="name.txt", mode="wb", fileobj=buf, mtime=0) as f:
        f.write(b"first member\n" * 100)
    m2 = gz(fastq(500, seed=4), 1)
    both = buf.getvalue() + m2 + b"\0" * 50
    want = b"first member\n" * 100 + gzip.decompress(m2)
    for piece in (1, 33, 1000, 1 << 20):
        assert gunzip(both, piece, len(want) + 5) == want
    big = fastq(12000, seed=5)
    for wbits in (25, 28, 31):  # 512-byte, 4 KiB and 32 KiB windows
        assert gunzip(gz(big, 9, wbits=wbits), 5000, len(big) + 1) == big


def test_truncation_corruption_and_capacity_are_errors():
    rng = random.Random(9)
    data = fastq(2000, seed=6)
    comp = gz(data, 6)
    for cut in (5, 12, len(comp) // 2, len(comp) - 9, len(comp) - 1):
        with pytest.raises(native.NativeError):
            gunzip(comp[:cut], 1000, 1 << 20)
    for _ in range(200):
        b = bytearray(comp)
        b[rng.randrange(10, len(b))] ^= 1 << rng.randrange(8)
        with pytest.raises(native.NativeError):
            gunzip(bytes(b), 1000, 1 << 21)
    with pytest.raises(native.NativeError):
        gunzip(comp, 1000, 100)
    with pytest.raises(native.NativeError):
        gunzip(b"", 1000, 100)
    with pytest.raises(native.NativeError):
        gunzip(b"not gzip at all", 1000, 100)


def test_text_reader_built_in_and_zlib_agree(tmp_path, monkeypatch):
    data = fastq(5000, seed=7)
    p = tmp_path / "r.fq.gz"
    p.write_bytes(gz(data[: len(data) // 3], 1) + gz(data[len(data) // 3:], 9))
    outs = []
    for use_zlib in ("0", "1"):
        monkeypatch.setenv("CSQ_ZLIB_INFLATE", use_zlib)
        got = b""
        with native.TextReader(str(p)) as r:
            while True:
                n, texts, _ = r.next(777)
                if n == 0:
                    break
                got += texts[0]
        outs.append(got)
    assert outs[0] == outs[1] == data


# ---- the parallel decoder of ordinary gzip files (csrc/pinflate.cpp) ----
@pytest.mark.parametrize("level", [1, 6, 9])
def test_parallel_inflate_matches_zlib(level):
    """One DEFLATE stream decoded by several threads: block starts searched behind every cut, spans decoded with
    markers for the unknown window, resolved in order.  Small spans make many spans and rounds out of a small file."""
    text = fastq(30000, seed=11)
    comp = gz(text, level)
    for span in (20_000, 150_000, 0):
        for threads in (2, 3, 8):
            assert native.pinflate(comp, len(text), threads, span) == text, (level, span, threads)


def test_parallel_inflate_members_stored_fixed_and_binary():
    text = fastq(8000, seed=12)
    # members that end in the middle of spans, empty and one-byte members, zero padding behind the last one
    parts = [text[:100], b"", text[100:300_000], text[300_000:300_001], text[300_001:]]
    comp = b"".join(gz(p) for p in parts) + bytes(64)
    for span in (5_000, 50_000, 1 << 20):
        assert native.pinflate(comp, len(text), 4, span) == text
    # stored and fixed blocks, tiny inputs
    for data in (text[:200_000], b"", b"A", b"ACGT\n" * 3):
        for level, strategy in ((0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_FIXED), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_HUFFMAN_ONLY)):
            assert native.pinflate(gz(data, level, strategy), len(data), 3, 10_000) == data
    # content that is not text: no block start passes the search, the first span's decoder walks the whole stream
    rng = random.Random(5)
    blob = os.urandom(100_000) + bytes(rng.choice(b"\x00\x01\xfe\xffAB") for _ in range(1_000_000))
    assert native.pinflate(gz(blob), len(blob), 4, 30_000) == blob


def test_parallel_inflate_reports_damage():
    text = fastq(20000, seed=13)
    comp = bytearray(gz(text))
    for pos in (30, len(comp) // 2, len(comp) - 6):
        bad = bytearray(comp)
        bad[pos] ^= 0x55
        with pytest.raises(native.NativeError):
            native.pinflate(bytes(bad), len(text) + 100_000, 4, 40_000)
    with pytest.raises(native.NativeError):
        native.pinflate(bytes(comp[: len(comp) // 2]), len(text), 4, 40_000)
    with pytest.raises(native.NativeError) as e:
        native.pinflate(bytes(comp), len(text) - 1, 4, 40_000)
    assert e.value.code == -5  # CSQ_ERR_CAPACITY


def test_text_reader_uses_the_parallel_decoder(tmp_path, monkeypatch):
    """The file driver's text reader hands large ordinary .gz inputs to the parallel decoder (threshold lowered here)."""
    recs = [fastq(3000, seed=21), fastq(3000, seed=22)]
    paths = [tmp_path / "a.fq.gz", tmp_path / "b.fq.gz"]
    for p, t in zip(paths, recs):
        p.write_bytes(gz(t[: len(t) // 2]) + gz(t[len(t) // 2:]))  # two members
    monkeypatch.setenv("CSQ_PINFLATE_MIN", "1000")
    monkeypatch.setenv("CSQ_PINFLATE_SPAN", "30000")
    monkeypatch.setenv("CSQ_INFLATE_THREADS", "4")
    got, total = [b"", b""], 0
    with native.TextReader(str(paths[0]), str(paths[1])) as r:
        while True:
            n, texts, _ = r.next(700)
            if n == 0:
                break
            got[0] += texts[0]
            got[1] += texts[1]
            total += n
    assert total == 3000 and got[0] == recs[0] and got[1] == recs[1]


def test_parallel_inflate_random_streams():
    """Random contents, members, levels, strategies, memory levels (tiny blocks), span sizes down to 1 KB and thread counts
    against zlib."""
    rng = random.Random(99)

    def reads(n):
        out = []
        for i in range(n):
            ln = rng.randint(1, 300)
            out.append(f"@M{rng.randint(0, 9)}:{i} {'x' * rng.randint(0, 30)}\n{''.join(rng.choice('ACGTN') for _ in range(ln))}\n+\n"
                       f"{''.join(rng.choice('FFFFFF:,#') for _ in range(ln))}\n")
        return "".join(out).encode()

    for it in range(40):
        kind = rng.choice(["fastq", "fastq", "mixed", "binary", "lines"])
        if kind == "fastq":
            data = reads(rng.randint(1, 4000))
        elif kind == "mixed":
            data = reads(rng.randint(100, 1500)) + os.urandom(rng.randint(0, 50000)) + reads(rng.randint(100, 1500))
        elif kind == "binary":
            data = os.urandom(rng.randint(0, 150000))
        else:
            data = b"".join(bytes(rng.choice(b"ab\n") for _ in range(rng.randint(0, 80))) + b"\n" for _ in range(rng.randint(0, 15000)))
        cuts = sorted(rng.randint(0, len(data)) for _ in range(rng.randint(0, 3)))
        parts = [data[a:b] for a, b in zip([0] + cuts, cuts + [len(data)])]
        comp = b""
        for p in parts:
            c = zlib.compressobj(rng.choice([1, 1, 4, 6, 6, 9]), zlib.DEFLATED, 31, rng.choice([1, 8, 9]),
                                 rng.choice([zlib.Z_DEFAULT_STRATEGY] * 4 + [zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FILTERED]))
            comp += c.compress(p) + c.flush()
        if rng.random() < 0.2:
            comp += bytes(rng.randint(1, 40))  # zero padding behind the last member
        span, threads = rng.choice([1000, 3000, 10000, 40000, 200000, 0]), rng.choice([2, 3, 4, 8, 16])
        assert native.pinflate(comp, len(data) + 1, threads, span) == data, (it, kind, len(data), span, threads)


def test_bench_fixture_writes_one_valid_gzip_member():
    """scripts/bench_files.gzip_single_member (the ordinary-gzip fixture of the bench): pieces deflated side by side and
    joined by sync flushes are ONE member that gzip, zlib and both of the library's decoders read."""
    from scripts import bench_files

    text = fastq(5000, seed=31)
    comp = bench_files.gzip_single_member(text, 4, level=6, piece=100_000)
    assert gzip.decompress(comp) == text
    d = zlib.decompressobj(31)
    assert d.decompress(comp) == text and d.eof and d.unused_data == b""  # a single member
    assert gunzip(comp, 1 << 20, len(text)) == text
    assert native.pinflate(comp, len(text), 4, 50_000) == text
