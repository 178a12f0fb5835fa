"""Library-scheme grammar and built-in scheme table (host side, API surface).

Mirrors the public names of the reference's ``cutseq/common.py`` so that code
written against cutseq keeps working:

* ``load_adapters`` / ``BUILDIN_ADAPTERS``   (reference common.py:15-33)
* ``reverse_complement``                     (reference common.py:36-46)
* ``remove_fq_suffix``                       (reference common.py:49-77)
* ``BarcodeSeq``                             (reference common.py:80-110)
* ``BarcodeConfig``                          (reference common.py:113-213)
* ``print_builtin_adapters``                 (reference common.py:216-235)

The scheme grammar is ``P5 [(INLINE5)] N* X* {>|<|-} X* N* [(INLINE3)] P7``
(reference common.py:173-176).  It is matched anchored at the start only, so
trailing characters after P7 are ignored exactly as ``re.match`` does there.
"""

from __future__ import annotations

import logging
import re
import sys
import textwrap
from importlib import resources

try:  # Python 3.11+
    import tomllib
except ImportError:  # pragma: no cover - older interpreters
    import tomli as tomllib


def load_adapters() -> dict:
    """Return ``{NAME: scheme}`` for every entry of ``adapters.toml`` that has a scheme."""
    text = resources.files(__package__).joinpath("adapters.toml").read_text(encoding="utf-8")
    table = tomllib.loads(text)
    return {name: entry["scheme"] for name, entry in table.items() if "scheme" in entry}


BUILDIN_ADAPTERS = load_adapters()

_COMPLEMENT = str.maketrans("ATGCatgc", "TACGtacg")


def reverse_complement(b: str) -> str:
    """Reverse complement; characters outside ``ATGCatgc`` are kept as they are."""
    return b[::-1].translate(_COMPLEMENT)


# Order matters: the first matching suffix wins (reference common.py:67-77 builds the
# list extension-major, mate-tag-minor).
_FQ_SUFFIXES = tuple(
    tag + "." + ext
    for ext in ("fastq.gz", "fq.gz", "fastq", "fq")
    for tag in ("_R1_001", "_R2_001", "_R1", "_R2", "")
)


def remove_fq_suffix(f: str) -> str:
    """Strip a trailing ``[_R1|_R2|_R1_001|_R2_001].{fastq,fq}[.gz]`` from a file name."""
    for suffix in _FQ_SUFFIXES:
        if f.endswith(suffix):
            return f[: len(f) - len(suffix)]
    return f


class BarcodeSeq:
    """A scheme component: forward text ``fw``, its reverse complement ``rc`` and ``len``."""

    __slots__ = ("fw", "rc", "len")

    def __init__(self, seq: str):
        self.fw = seq
        self.rc = reverse_complement(seq)
        self.len = len(seq)

    def __repr__(self) -> str:
        return f"{self.fw} ({self.rc})" if self.len else ""


_BASES = "[ATGCatgc]+"
_SCHEME_RE = re.compile(
    rf"(?P<p5>{_BASES})"
    rf"(?:\((?P<inline5>{_BASES})\))?"
    r"(?P<umi5>N*)(?P<mask5>X*)"
    r"(?P<strand>[-<>])"
    r"(?P<mask3>X*)(?P<umi3>N*)"
    rf"(?:\((?P<inline3>{_BASES})\))?"
    rf"(?P<p7>{_BASES})"
)

_PARTS = ("p5", "p7", "inline5", "inline3", "umi5", "umi3", "mask5", "mask3")
_STRAND = {">": "+", "<": "-", "-": None}


class BarcodeConfig:
    """Parsed library scheme.

    Attributes ``p5 p7 inline5 inline3 umi5 umi3 mask5 mask3`` are ``BarcodeSeq``;
    ``strand`` is ``"+"`` (``>``), ``"-"`` (``<``) or ``None`` (``-``).
    An unparsable scheme logs an error and exits with status 1 like the reference
    (common.py:177-179).
    """

    def __init__(self, adapter: str | None = None):
        self.strand = None
        for part in _PARTS:
            setattr(self, part, BarcodeSeq(""))
        if adapter is not None:
            self._parse_barcode(adapter)

    def _parse_barcode(self, b: str) -> None:
        m = _SCHEME_RE.match(b)
        if m is None:
            logging.error(f"barcode {b} is not valid")
            sys.exit(1)
        self.strand = _STRAND[m.group("strand")]
        for part in _PARTS:
            setattr(self, part, BarcodeSeq(m.group(part) or ""))

    def to_dict(self) -> dict:
        d = {part: getattr(self, part).fw for part in _PARTS}
        d["strand"] = self.strand
        return d


def print_builtin_adapters() -> None:
    """``--list-adapters``: aligned two-column table, schemes wrapped at 100 columns."""
    name_w = max(map(len, BUILDIN_ADAPTERS))
    scheme_w = max(map(len, BUILDIN_ADAPTERS.values()))
    print("\nBuilt-in adapter schemes:\n")
    print(f"{'Name'.ljust(name_w)}   Scheme")
    print(f"{'-' * name_w}   {'-' * max(30, min(scheme_w, 100))}")
    for name, scheme in BUILDIN_ADAPTERS.items():
        lines = textwrap.wrap(scheme, width=100)
        print(f"{name.ljust(name_w)}   {lines[0]}")
        for extra in lines[1:]:
            print(f"{' ' * name_w}   {extra}")
    print("\nUse the adapter name with -A/--adapter-name, or the scheme string with -a/--adapter-scheme.\n")
