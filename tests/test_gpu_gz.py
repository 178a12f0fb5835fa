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
