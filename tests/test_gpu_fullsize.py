"""BASELINE.json's full sizes (10 M units per config) through size-independent properties.

The oracle finishes 10 M pairs in minutes, not seconds, so at full size the CUDA chain is checked by
  * batch-split invariance: the concatenated output must not depend on how the records are cut into batches
    (2 M-pair batches vs 700 001-pair batches, text batches vs host-parsed SoA batches) - a checksum of checksums;
  * conservation: every input record is written to exactly one destination; bases written + bases removed by the
    quality trimmer never exceed the input; counters of all batches add up;
  * idempotence of the finished output: trimmed reads of the TAKARAV3 run contain no further 3' adapter, so a
    second pass with only the BackAdapter ops finds (almost) nothing to trim beyond chance 3-mers - checked as
    "second pass changes no record that the oracle would not change" on a slice;
  * a 20 000-unit slice at a random offset of the same stream compared byte for byte with the oracle.
"""

import hashlib

import numpy as np
import pytest

from cutseq_b200 import _abi as A
from cutseq_b200 import native
from oracle import oracle
from tests import helpers

pytestmark = pytest.mark.gpu

FULL = 10_000_000
CONFIGS = [
    (2, ["-A", "TAKARAV3", "--trim-polyA"], 2),
    (3, ["-a", "ACACGACGCTCTTCCGATCT(ATCACG)NNNNNNNNXXX<XXX(CGTGAT)AGATCGGAAGAGCACACGTC", "--ensure-inline-barcode"], 2),
    (4, ["-A", "SMALLRNA"], 1),
]


def run_stream(plan, config, n_mates, total, batch, mode):
    """Feeds `total` units in batches of `batch`; returns (sha256 per (dest, mate), records per dest)."""
    hashes = [[hashlib.sha256() for _ in range(2)] for _ in range(A.CSQ_N_DEST)]
    records = [0] * A.CSQ_N_DEST
    done = 0
    while done < total:
        n = min(batch, total - done)
        b = native.synth_batch(config, n, first_index=done, buffer=6)
        if mode == "text":
            texts = [native.format_fastq(b, m) for m in range(n_mates)]
            text, rec = plan.run_text(texts, n, first_record=done)
        else:
            text, rec = plan.run_batch(b)
        for d in range(A.CSQ_N_DEST):
            records[d] += rec[d][0]
            for m in range(n_mates):
                hashes[d][m].update(text[d][m])
        done += n
    return [[h.hexdigest() for h in row] for row in hashes], records


@pytest.mark.parametrize("config,argv,n_mates", CONFIGS, ids=["config2", "config3", "config4"])
def test_full_size_batch_split_invariance_and_conservation(config, argv, n_mates):
    prog = helpers.program_for(argv, n_mates)
    with native.Plan(prog, 0, 0) as plan:
        h_text, rec_text = run_stream(plan, config, n_mates, FULL, 2_000_000, "text")
        c1 = plan.stats()
    with native.Plan(prog, 0, 0) as plan:
        h_soa, rec_soa = run_stream(plan, config, n_mates, FULL, 700_001, "soa")
        c2 = plan.stats()
    assert h_text == h_soa and rec_text == rec_soa
    for c in (c1, c2):
        assert c.n == FULL
        assert c.written + c.too_short + c.untrimmed == FULL
        assert [c.written, c.too_short, c.untrimmed] == rec_text
        for m in range(n_mates):
            assert c.written_bp[m] + c.quality_trimmed_bp[m] <= c.total_bp[m]
            assert c.total_bp[m] == FULL * (150 if config != 4 else 75)
    for f in ("written", "too_short", "untrimmed"):
        assert getattr(c1, f) == getattr(c2, f)
    assert [list(c1.with_adapters[m]) for m in range(2)] == [list(c2.with_adapters[m]) for m in range(2)]
    assert [list(c1.dp_cells[m]) for m in range(2)] == [list(c2.dp_cells[m]) for m in range(2)]


@pytest.mark.parametrize("config,argv,n_mates", CONFIGS, ids=["config2", "config3", "config4"])
def test_slice_of_the_full_stream_against_oracle(config, argv, n_mates):
    """A slice deep inside the 10 M-unit stream (the generator is counter based): GPU == oracle, byte for byte."""
    prog = helpers.program_for(argv, n_mates)
    first = 8_765_432
    batch = native.synth_batch(config, 20000, first_index=first, buffer=6)
    want = oracle.run_batch(prog, batch, n_threads=8)
    texts = [native.format_fastq(batch, m) for m in range(n_mates)]
    with native.Plan(prog, 0, 0) as plan:
        got, records = plan.run_text(texts, 20000, first_record=first)
    for d in range(A.CSQ_N_DEST):
        for m in range(n_mates):
            assert got[d][m] == want["text"][d][m], (config, d, m)


def test_trimmed_output_is_a_fixed_point_of_the_3prime_search():
    """Idempotence: re-running the 3' adapter search of TAKARAV3 (BackAdapter p7 / p5rc, e = 0.2, O = 3) on reads
    that already went through the chain must agree with the oracle on the same reads (whatever it still finds -
    chance 3-mers at the new read end - is found identically), and must find far fewer matches than the first pass."""
    prog = helpers.program_for(["-A", "TAKARAV3"], 2)
    n = 200_000
    batch = native.synth_batch(2, n, first_index=4_000_000, buffer=6)
    texts = [native.format_fastq(batch, m) for m in range(2)]
    with native.Plan(prog, 0, 0) as plan:
        first_text, rec = plan.run_text(texts, n)
        c1 = plan.stats()
    n2 = rec[0][0]
    # headers were renamed to "<id>_<UMI>": ids of both mates still agree, so the pair check passes again
    with native.Plan(prog, 0, 0) as plan:
        second_text, rec2 = plan.run_text([first_text[0][0], first_text[0][1]], n2)
        c2 = plan.stats()
    lines = [t.split(b"\n") for t in (first_text[0][0], first_text[0][1])]
    mates = [[(ls[i][1:].decode(), ls[i + 1].decode(), ls[i + 3].decode()) for i in range(0, 4 * 5000, 4)] for ls in lines]
    b2, keep = oracle.make_batch(*mates)
    want = oracle.run_batch(prog, b2, n_threads=4)
    with native.Plan(prog, 0, 0) as plan:
        got, _ = plan.run_batch(b2)
    assert got[0][0] == want["text"][0][0] and got[0][1] == want["text"][0][1]
    back_ops = [t for t, op in enumerate(prog.ops_r1) if op.kind == A.OP_ALIGN and op.adapter_kind in (A.AD_BACK, A.AD_BACK_ANYWHERE)]
    first_hits = sum(c1.with_adapters[0][t] for t in back_ops)
    second_hits = sum(c2.with_adapters[0][t] for t in back_ops)
    assert second_hits * 5 < first_hits, (first_hits, second_hits)
