"""TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/liboracle.so (cutseq_oracle.c) plus
small helpers to build SoA batches from Python records.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by
cutseq_b200/ (the product path has no CPU fallback)."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from cutseq_b200 import _abi as A

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class orc_match(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("found", "ref_start", "ref_stop", "query_start", "query_stop", "score", "errors")]


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "cutseq_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "cutseq_b200.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "liboracle.so"])
    ssrc = os.path.join(HERE, "..", "cutseq_b200", "csrc", "synth.cu")
    sso = os.path.join(HERE, "libsynth.so")
    if force or not os.path.exists(sso) or os.path.getmtime(sso) < os.path.getmtime(ssrc):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "libsynth.so"])
    return so


_SYNTH = None


def synth_batch(config: int, n_reads: int, first_index: int = 0, buffer: int = 0) -> A.csq_batch_in:
    """bench.py's synthetic workload (cutseq_b200/csrc/synth.cu) from oracle/libsynth.so: the same generator built
    without CUDA, for the CPU arm (which must not map the product library)."""
    global _SYNTH
    if _SYNTH is None:
        so = os.path.join(HERE, "libsynth.so")
        src = os.path.join(HERE, "..", "cutseq_b200", "csrc", "synth.cu")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", HERE, "-B", "libsynth.so"])
        _SYNTH = C.CDLL(so)
        _SYNTH.csq_synth_batch.argtypes = [C.POINTER(A.csq_synth), C.c_uint64, C.c_uint32, C.c_int, C.POINTER(A.csq_batch_in)]
    defaults = {2: (20240419, 150, True), 3: (20240420, 150, True), 4: (20240421, 75, False), 5: (20240419, 150, True)}
    seed, read_len, paired = defaults[config]
    cfg = A.csq_synth(seed, read_len, int(paired), 2 if config == 5 else config)
    b = A.csq_batch_in()
    if _SYNTH.csq_synth_batch(C.byref(cfg), first_index, n_reads, buffer, C.byref(b)) != 0:
        raise RuntimeError("synthetic generator failed")
    return b


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_locate.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(orc_match)]
        _LIB.orc_adapter_match.argtypes = [C.POINTER(A.csq_op), C.c_char_p, C.c_int, C.POINTER(orc_match)]
        _LIB.orc_quality_trim_index.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _LIB.orc_run_batch.argtypes = [
            C.POINTER(A.csq_op), C.c_int, C.POINTER(A.csq_op), C.c_int, C.POINTER(A.csq_filters),
            C.POINTER(A.csq_batch_in), C.POINTER(A.csq_batch_out), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.POINTER(A.csq_counters), C.c_int,
        ]
        _LIB.orc_free.argtypes = [C.POINTER(A.csq_batch_out)]
        _LIB.orc_max_threads.restype = C.c_int
        _LIB.orc_names_match.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    return _LIB


def locate(reference: str, query: str, max_error_rate: float, flags: int, min_overlap: int = 1):
    """Aligner(reference, max_error_rate, flags, min_overlap=..).locate(query)"""
    m = orc_match()
    found = lib().orc_locate(reference.encode(), len(reference), max_error_rate, flags, min_overlap, query.encode(), len(query), C.byref(m))
    if not found:
        return None
    return (m.ref_start, m.ref_stop, m.query_start, m.query_stop, m.score, m.errors)


def adapter_match(op, query: str):
    cop = op.to_c() if hasattr(op, "to_c") else op
    m = orc_match()
    found = lib().orc_adapter_match(C.byref(cop), query.encode(), len(query), C.byref(m))
    if not found:
        return None
    return (m.ref_start, m.ref_stop, m.query_start, m.query_stop, m.score, m.errors)


def names_match(header1: str, header2: str) -> bool:
    """dnaio.record_names_match(header1, header2)"""
    a, b = header1.encode("latin-1"), header2.encode("latin-1")
    return bool(lib().orc_names_match(a, len(a), b, len(b)))


def quality_trim_index(qualities: str, cutoff_front: int, cutoff_back: int, base: int = 33):
    a, b = C.c_int(), C.c_int()
    lib().orc_quality_trim_index(qualities.encode(), len(qualities), cutoff_front, cutoff_back, base, C.byref(a), C.byref(b))
    return (a.value, b.value)


class Mate:
    """Packed SoA of one mate (layout of csq_mate_in), backed by numpy arrays."""

    def __init__(self, records):
        n = len(records)
        lens = np.fromiter((len(r[1]) for r in records), dtype=np.uint32, count=n)
        padded = (lens.astype(np.uint64) + 15) // 16 * 16
        off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(padded, out=off[1:])
        total = int(off[-1]) + 16
        self.seq = np.full(total, 0, dtype=np.uint8)
        self.qual = np.full(total, 0, dtype=np.uint8)
        self.seq_off = off[:-1].astype(np.uint32)
        self.seq_len = lens
        names = [r[0].encode("latin-1") if isinstance(r[0], str) else r[0] for r in records]
        noff = np.zeros(n + 1, dtype=np.uint32)
        if n:
            np.cumsum([len(x) for x in names], out=noff[1:])
        self.name = np.frombuffer(b"".join(names) + b"\0", dtype=np.uint8).copy()
        self.name_off = noff
        for i, r in enumerate(records):
            s = r[1].encode("latin-1") if isinstance(r[1], str) else r[1]
            q = r[2].encode("latin-1") if isinstance(r[2], str) else r[2]
            o = int(off[i])
            self.seq[o : o + len(s)] = np.frombuffer(s, dtype=np.uint8)
            self.qual[o : o + len(q)] = np.frombuffer(q, dtype=np.uint8)
        self.n = n

    def to_c(self) -> A.csq_mate_in:
        m = A.csq_mate_in()
        m.seq = self.seq.ctypes.data
        m.qual = self.qual.ctypes.data
        m.seq_off = self.seq_off.ctypes.data
        m.seq_len = self.seq_len.ctypes.data
        m.seq_bytes = self.seq.size // 16 * 16
        m.name = self.name.ctypes.data
        m.name_off = self.name_off.ctypes.data
        m.name_bytes = int(self.name_off[-1]) if self.n else 0
        return m


def make_batch(records1, records2=None):
    """records: list of (name, seq, qual). Returns (csq_batch_in, keepalive)."""
    m1 = Mate(records1)
    b = A.csq_batch_in()
    b.n_reads = m1.n
    b.n_mates = 1
    b.mate[0] = m1.to_c()
    keep = [m1]
    if records2 is not None:
        m2 = Mate(records2)
        assert m2.n == m1.n
        b.n_mates = 2
        b.mate[1] = m2.to_c()
        keep.append(m2)
    return b, keep


def parse_fastq_text(text: bytes):
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    assert len(lines) % 4 == 0
    return [(lines[i][1:].decode("latin-1"), lines[i + 1].decode("latin-1"), lines[i + 3].decode("latin-1")) for i in range(0, len(lines), 4)]


def run_batch(program, batch, n_threads: int = 1, want_matches: bool = True):
    """Run the whole chain on the CPU oracle. Returns dict(text[d][m] bytes, results, matches, counters, status)."""
    ops1, n1 = program.c_ops(0)
    ops2, n2 = program.c_ops(1)
    flt = program.filters.to_c()
    out = A.csq_batch_out()
    n = batch.n_reads
    paired = batch.n_mates == 2
    res = [np.zeros(max(n, 1), dtype=np.dtype([("start", "u4"), ("stop", "u4"), ("dest", "u4"), ("matched", "u4")])) for _ in range(2)]
    mt_dtype = np.dtype([(k, "i2") for k in ("found", "ref_start", "ref_stop", "query_start", "query_stop", "score", "errors", "reserved")])
    mts = [np.zeros((max(n1, 1), max(n, 1)), dtype=mt_dtype), np.zeros((max(n2, 1), max(n, 1)), dtype=mt_dtype)]
    counters = A.csq_counters()
    status = lib().orc_run_batch(
        ops1, n1, ops2, n2 if paired else 0, C.byref(flt), C.byref(batch), C.byref(out),
        res[0].ctypes.data, res[1].ctypes.data if paired else None,
        mts[0].ctypes.data if want_matches else None, mts[1].ctypes.data if (want_matches and paired) else None,
        C.byref(counters), n_threads,
    )
    text = [[C.string_at(out.text[d][m].data, out.text[d][m].bytes) if out.text[d][m].data else b"" for m in range(2)] for d in range(A.CSQ_N_DEST)]
    records = [[out.text[d][m].records for m in range(2)] for d in range(A.CSQ_N_DEST)]
    lib().orc_free(C.byref(out))
    return {"text": text, "records": records, "results": [r[:n] for r in res], "matches": [m[:, :n] for m in mts], "counters": counters, "status": status}
