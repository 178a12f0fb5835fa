"""Device gzip writer (CSQ_PLAN_GZIP_OUT): every output stream arrives as concatenated BGZF members whose decompressed
bytes are the FASTQ text the plain plan delivers (= the oracle's, tests/test_gpu_parity.py)."""

import struct
import zlib

import pytest

from cutseq_b200 import _abi as A
from cutseq_b200 import native
from oracle import oracle
from tests import helpers

pytestmark = pytest.mark.gpu


def gunzip_members(data, max_isize=256 * 127):
    out = []
    while data:
        assert data[:4] == b"\x1f\x8b\x08\x04" and data[12:16] == b"BC\x02\x00"
        size = struct.unpack_from("<H", data, 16)[0] + 1
        d = zlib.decompressobj(31)
        piece = d.decompress(data[:size])
        assert d.eof and not d.unused_data and len(piece) <= max_isize
        out.append(piece)
        data = data[size:]
    return b"".join(out)


@pytest.mark.parametrize("config,argv,n_mates,n", [
    (2, ["-A", "TAKARAV3", "--trim-polyA"], 2, 60000),
    (3, ["-a", "ACACGACGCTCTTCCGATCT(ATCACG)NNNNNNNNXXX<XXX(CGTGAT)AGATCGGAAGAGCACACGTC", "--ensure-inline-barcode"], 2, 20000),
    (4, ["-A", "SMALLRNA"], 1, 50000),
    (2, ["-A", "TAKARAV3"], 2, 1),
    (2, ["-A", "TAKARAV3"], 2, 0),
])
def test_gzip_streams_hold_the_text(config, argv, n_mates, n):
    prog = helpers.program_for(argv, n_mates)
    batch = native.synth_batch(config, n, first_index=4242, buffer=4)
    with native.Plan(prog, 0, 0) as plan:
        text, records = plan.run_batch(batch)
    with native.Plan(prog, 0, A.PLAN_GZIP_OUT) as plan:
        for rep in range(2):  # the second batch reuses the slot's buffers
            z, zrecords = plan.run_batch(batch)
            assert zrecords == records
            for d in range(A.CSQ_N_DEST):
                for m in range(n_mates):
                    assert gunzip_members(z[d][m]) == text[d][m], (d, m)
                    if text[d][m]:
                        assert len(z[d][m]) < 0.6 * len(text[d][m]) + 200


def test_small_output_buffers_can_be_retried():
    """CSQ_ERR_CAPACITY reports the packed sizes; a second csq_wait with larger buffers fetches the same members."""
    import numpy as np

    prog = helpers.program_for(["-A", "TAKARAV3", "--trim-polyA"], 2)
    batch = native.synth_batch(2, 30000, first_index=1, buffer=4)
    with native.Plan(prog, 0, 0) as plan:
        text, _ = plan.run_batch(batch)
    with native.Plan(prog, 0, A.PLAN_GZIP_OUT) as plan:
        out = A.csq_batch_out()
        small = [[np.empty(1000, dtype=np.uint8) for _ in range(2)] for _ in range(A.CSQ_N_DEST)]
        for d in range(A.CSQ_N_DEST):
            for m in range(2):
                out.text[d][m].data = small[d][m].ctypes.data
                out.text[d][m].capacity = 1000
        plan.submit(0, batch, out)
        with pytest.raises(native.NativeError) as e:
            plan.wait(0)
        assert e.value.code == A.ERR_CAPACITY
        big = [[np.empty(int(out.text[d][m].bytes) + 16, dtype=np.uint8) for m in range(2)] for d in range(A.CSQ_N_DEST)]
        for d in range(A.CSQ_N_DEST):
            for m in range(2):
                out.text[d][m].data = big[d][m].ctypes.data
                out.text[d][m].capacity = big[d][m].size
        plan.wait(0)
        for d in range(A.CSQ_N_DEST):
            for m in range(2):
                assert gunzip_members(big[d][m][: out.text[d][m].bytes].tobytes()) == text[d][m]


# ---- BGZF batches in: the device inflates whole members and finds the records itself ----
def _bgzf(text, level=6, piece=0xFF00):
    from tests.test_gz import bgzf

    return bgzf(text, level, zlib.Z_DEFAULT_STRATEGY, piece)


def _fastq(batch, m):
    return bytes(native.format_fastq(batch, m))


@pytest.mark.parametrize("level,piece", [(1, 0xFF00), (6, 0xFF00), (6, 5000), (0, 0xFF00)])
def test_bgzf_batches_match_text_batches(level, piece):
    prog = helpers.program_for(["-A", "TAKARAV3", "--trim-polyA"], 2)
    n = 30000
    batch = native.synth_batch(2, n, first_index=99, buffer=5)
    texts = [_fastq(batch, m) for m in range(2)]
    with native.Plan(prog, 0, 0) as plan:
        want, wrec = plan.run_text(texts, n)
        runs = [native.BgzfRun(_bgzf(t, level, piece)) for t in texts]
        # pass 1 of the file driver: line ends per member
        for m in range(2):
            lines = plan.bgzf_count_lines(runs[m])
            assert int(lines.sum()) == texts[m].count(b"\n") == 4 * n
            assert lines[0] == texts[m][: min(piece, len(texts[m]))].count(b"\n")
        got, rec = plan.run_bgzf(runs, n, capacity=len(texts[0]) + 64 * n)
        assert got == want and rec == wrec
        # a batch in the middle of the members: 1000 records in front of it, whatever follows is ignored
        sub = 5000
        wsub, _ = plan.run_text([b"\n".join(t.split(b"\n")[4000 : 4000 + 4 * sub]) + b"\n" for t in texts], sub)
        runs2 = [native.BgzfRun(_bgzf(t, level, piece), skip_lines=4000) for t in texts]
        gsub, _ = plan.run_bgzf(runs2, sub, capacity=len(texts[0]))
        assert gsub == wsub


@pytest.mark.parametrize("strategy", [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY], ids=["default", "fixed", "rle", "huffman"])
def test_device_inflate_on_every_kind_of_stream(strategy):
    """Content that drives every path of the warp decoder: stored, fixed and dynamic blocks, long matches copied by all
    lanes, runs (distance 1), short periods (distance < length), long codes behind the direct tables, literal-only
    streams.  k_gz_check compares the CRC-32 of what was inflated with every member's trailer, so a wrong byte anywhere
    is an error; the line ends per member are compared with the text."""
    import os
    import random

    from tests.test_gz import bgzf

    rng = random.Random(17)
    contents = {
        "fastq": _fastq(native.synth_batch(2, 3000, first_index=7, buffer=5), 0),
        "runs": b"".join(bytes([rng.choice(b"ACGT\n")]) * rng.randint(1, 700) for _ in range(2000)),
        "periods": b"".join((bytes(rng.choice(b"ACGT#I\n") for _ in range(p)) * (3000 // p)) for p in (2, 3, 5, 7, 8, 9, 15, 16, 17, 31, 33, 100)),
        "random": os.urandom(200_000),
        "skewed": bytes(rng.choice(b"A" * 200 + bytes(range(256))) for _ in range(300_000)),  # code lengths up to 15
        "one": b"\n",
    }
    prog = helpers.program_for(["-A", "SMALLRNA"], 1)
    with native.Plan(prog, 0, 0) as plan:
        for name, text in contents.items():
            for level in (1, 9):
                for piece in (0xFF00, 4093):
                    if not text:
                        continue
                    z = bgzf(text, level, strategy, piece)
                    lines = plan.bgzf_count_lines(native.BgzfRun(z))
                    want = [text[o:o + piece].count(b"\n") for o in range(0, len(text), piece)]
                    assert [int(v) for v in lines] == want, (name, level, piece)


def test_bgzf_in_gzip_out_round_trip_and_corrupt_member():
    prog = helpers.program_for(["-A", "TAKARAV3"], 2)
    n = 20000
    batch = native.synth_batch(2, n, first_index=5, buffer=5)
    texts = [_fastq(batch, m) for m in range(2)]
    with native.Plan(prog, 0, 0) as plan:
        want, _ = plan.run_text(texts, n)
    with native.Plan(prog, 0, A.PLAN_GZIP_OUT) as plan:
        z, _ = plan.run_bgzf([native.BgzfRun(_bgzf(t, 1)) for t in texts], n, capacity=len(texts[0]))
        for d in range(A.CSQ_N_DEST):
            for m in range(2):
                assert gunzip_members(z[d][m]) == want[d][m]
        # the library reads its own output: trimmed files as the next run's input
        again = [native.BgzfRun(z[0][m]) for m in range(2)]
        assert int(plan.bgzf_count_lines(again[0]).sum()) == want[0][0].count(b"\n")
        bad = bytearray(_bgzf(texts[0], 6))
        bad[len(bad) // 3] ^= 0xFF
        bad[len(bad) // 3 + 1] ^= 0xFF
        with pytest.raises(native.NativeError) as e:
            plan.run_bgzf([native.BgzfRun(bytes(bad)), native.BgzfRun(_bgzf(texts[1], 6))], n, capacity=len(texts[0]))
        assert e.value.code in (A.ERR_IO, A.ERR_FORMAT)
