"""Minimal stand-in for dnaio.SequenceRecord (upstream dnaio src/dnaio/_core.pyx) and a
4-line FASTQ reader/writer. Test infrastructure only."""

import gzip
import io

_COMP = bytes.maketrans(b"ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", b"TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn").decode("latin-1")


class SequenceRecord:
    __slots__ = ("name", "sequence", "qualities")

    def __init__(self, name, sequence, qualities=None):
        if qualities is not None and len(qualities) != len(sequence):
            raise ValueError("size of sequence and qualities differ")
        self.name, self.sequence, self.qualities = name, sequence, qualities

    def __getitem__(self, key):
        return SequenceRecord(self.name, self.sequence[key],
                              self.qualities[key] if self.qualities is not None else None)

    def __len__(self):
        return len(self.sequence)

    @property
    def id(self):
        return self.name.split(maxsplit=1)[0] if self.name.split(maxsplit=1) else self.name

    def reverse_complement(self):
        seq = self.sequence[::-1].translate({ord(a): b for a, b in zip(
            "ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn")})
        qual = self.qualities[::-1] if self.qualities is not None else None
        return SequenceRecord(self.name, seq, qual)

    def fastq_bytes(self):
        return f"@{self.name}\n{self.sequence}\n+\n{self.qualities}\n".encode("ascii")


def record_names_match(header1: str, header2: str) -> bool:
    """dnaio.record_names_match / SequenceRecord.is_mate (upstream src/dnaio/_core.pyx record_ids_match): header 2's id
    ends at its first space or tab; header 1 must end or carry a space / tab there; a trailing 1, 2 or 3 on BOTH ids is
    not compared."""
    id2 = len(header2)
    for i, ch in enumerate(header2):
        if ch in " \t":
            id2 = i
            break
    if len(header1) < id2:
        return False
    if id2 < len(header1) and header1[id2] not in " \t":
        return False
    if id2 > 0 and header1[id2 - 1] in "123" and header2[id2 - 1] in "123":
        id2 -= 1
    return header1[:id2] == header2[:id2]


def open_maybe_gz(path, mode):
    if str(path).endswith(".gz"):
        return gzip.open(path, mode, compresslevel=1) if "w" in mode else gzip.open(path, mode)
    return open(path, mode)


def read_fastq(path):
    with open_maybe_gz(path, "rb") as f:
        data = f.read()
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    if len(lines) % 4:
        raise ValueError("FASTQ file ended prematurely")
    for i in range(0, len(lines), 4):
        h, s, p, q = lines[i : i + 4]
        if not h.startswith(b"@") or not p.startswith(b"+"):
            raise ValueError(f"malformed FASTQ record at line {i + 1}")
        yield SequenceRecord(h[1:].rstrip(b"\r").decode("ascii"), s.rstrip(b"\r").decode("ascii"),
                             q.rstrip(b"\r").decode("ascii"))
