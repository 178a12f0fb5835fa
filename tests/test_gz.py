"""The device gzip writer's algorithm, run by its host twin (csq_gz_deflate_host): code construction, block header,
bit packing and CRC folding produce valid BGZF members that zlib / gzip read back byte for byte."""

import random
import struct
import subprocess
import zlib

import pytest

from cutseq_b200 import native

PIECE = 256 * 127


def fastq_text(rng, n_bytes):
    recs = []
    size = 0
    i = 0
    while size < n_bytes:
        L = rng.randint(30, 150)
        seq = "".join(rng.choice("ACGTN" if rng.random() < 0.02 else "ACGT") for _ in range(L))
        qual = "".join(rng.choice("I9-#") for _ in range(L))
        r = f"@SIM:1:FC:1:{1101 + i // 1000}:{rng.randint(1000, 30000)}:{rng.randint(1000, 30000)}_{''.join(rng.choice('ACGT') for _ in range(8))}\n{seq}\n+\n{qual}\n"
        recs.append(r)
        size += len(r)
        i += 1
    return "".join(recs).encode()[:n_bytes]


def gunzip_members(data):
    out, members = [], 0
    while data:
        d = zlib.decompressobj(31)
        out.append(d.decompress(data))
        assert d.eof
        data = d.unused_data
        members += 1
    return b"".join(out), members


@pytest.mark.parametrize("n", [0, 1, 2, 100, PIECE - 1, PIECE, PIECE + 1, 3 * PIECE, 3 * PIECE + 12345, 700_000])
def test_members_decode_to_the_text(n):
    rng = random.Random(n)
    text = fastq_text(rng, n)
    z = native.gz_deflate_host(text)
    back, members = gunzip_members(z)
    assert back == text
    assert members == (n + PIECE - 1) // PIECE
    # BGZF framing: the 'BC' field holds the member size - 1; CRC-32 and ISIZE close the member
    pos = 0
    off = 0
    while pos < len(z):
        assert z[pos : pos + 4] == b"\x1f\x8b\x08\x04" and z[pos + 12 : pos + 16] == b"BC\x02\x00"
        size = struct.unpack_from("<H", z, pos + 16)[0] + 1
        crc, isize = struct.unpack_from("<II", z, pos + size - 8)
        assert isize == min(PIECE, n - off) and crc == zlib.crc32(text[off : off + isize])
        pos += size
        off += isize
    assert pos == len(z) and off == n


def test_binary_and_incompressible_pieces_are_stored():
    rng = random.Random(7)
    noise = bytes(rng.getrandbits(8) for _ in range(2 * PIECE + 999))
    z = native.gz_deflate_host(noise)
    back, members = gunzip_members(z)
    assert back == noise and members == 3
    assert len(z) <= len(noise) + 3 * (18 + 8 + 5)
    every = bytes(range(256)) * 300  # all 256 literals used
    assert gunzip_members(native.gz_deflate_host(every))[0] == every
    one = b"A" * (PIECE + 5)
    assert gunzip_members(native.gz_deflate_host(one))[0] == one


def test_size_is_the_order0_huffman_cost_and_gzip_accepts_the_file(tmp_path):
    import collections
    import heapq

    rng = random.Random(11)
    text = fastq_text(rng, 1_500_000)
    z = native.gz_deflate_host(text)
    # cost of an optimal order-0 Huffman code of the whole text (sum of the merged weights)
    heap = list(collections.Counter(text).values())
    heapq.heapify(heap)
    cost = 0
    while len(heap) > 1:
        a, b = heapq.heappop(heap), heapq.heappop(heap)
        cost += a + b
        heapq.heappush(heap, a + b)
    assert len(z) < 1.03 * cost / 8            # sampled histogram, all 256 literals coded, header + framing per member
    assert len(z) < 1.30 * len(zlib.compress(text, 1))  # literal-only Huffman against level-1 LZ77 + Huffman (random qualities here)
    p = tmp_path / "x.fastq.gz"
    p.write_bytes(z)
    assert subprocess.run(["gzip", "-t", str(p)]).returncode == 0
    assert subprocess.run(["gzip", "-dc", str(p)], capture_output=True).stdout == text


# ---- the device gzip reader's decoder (gz_inflate_core.h), run by its host twin ----
def bgzf(text, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, piece=0xFF00):
    out = []
    for o in range(0, len(text), piece):
        p = text[o : o + piece]
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        body = c.compress(p) + c.flush()
        size = 18 + len(body) + 8
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", size - 1) + body
                   + struct.pack("<II", zlib.crc32(p) & 0xFFFFFFFF, len(p)))
    return b"".join(out)


@pytest.mark.parametrize("level,strategy", [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY),
                                            (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)])
def test_inflate_reads_bgzf_members(level, strategy):
    rng = random.Random(level * 10 + strategy)
    for n in (1, 17, 5000, 0xFF00, 0xFF00 + 1, 300_000):
        text = fastq_text(rng, n)
        z = bgzf(text, level, strategy)
        back, lines = native.gz_inflate_host(z, len(text))
        assert back == text and lines == text.count(b"\n"), (n, level, strategy)
    # the BGZF end-of-file marker (an empty member) and members in a row
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    text = fastq_text(rng, 100_000)
    back, lines = native.gz_inflate_host(bgzf(text, level, strategy) + eof, len(text))
    assert back == text


def test_inflate_reads_what_the_writer_produces_and_rejects_damage():
    rng = random.Random(3)
    text = fastq_text(rng, 200_000) + bytes(rng.getrandbits(8) for _ in range(40_000))  # the tail is stored
    z = native.gz_deflate_host(text)
    back, lines = native.gz_inflate_host(z, len(text))
    assert back == text and lines == text.count(b"\n")
    noise = bytes(rng.getrandbits(8) for _ in range(70_000))
    assert native.gz_inflate_host(bgzf(noise, 6), len(noise))[0] == noise
    long_runs = b"A" * 50_000 + b"\n" + b"ACGT" * 5_000  # distance-1 and short-period matches
    assert native.gz_inflate_host(bgzf(long_runs, 9), len(long_runs))[0] == long_runs
    bad = bytearray(bgzf(fastq_text(rng, 30_000), 6))
    bad[len(bad) // 2] ^= 0x55
    with pytest.raises(native.NativeError):
        native.gz_inflate_host(bytes(bad), 30_000)
    with pytest.raises(native.NativeError):
        native.gz_inflate_host(b"\x1f\x8b\x08\x00" + b"\x00" * 30, 100)  # a gzip member without the BGZF field
