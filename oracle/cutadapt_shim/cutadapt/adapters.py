"""Restatement of the adapter classes cutseq uses (upstream src/cutadapt/adapters.py:
``Where``, ``SingleAdapter``, ``FrontAdapter``, ``RightmostFrontAdapter``, ``BackAdapter``,
``NonInternalFrontAdapter``, ``NonInternalBackAdapter``, ``PrefixAdapter``,
``SuffixAdapter``, ``RemoveBeforeMatch``, ``RemoveAfterMatch``, ``MultipleAdapters``).
The k-mer heuristic in front of ``Aligner.locate`` is lossless upstream and is omitted.
"""

from enum import IntFlag

from .align import Aligner, EndSkip


class Where(IntFlag):
    BACK = EndSkip.QUERY_START | EndSkip.QUERY_STOP | EndSkip.REFERENCE_END
    FRONT = EndSkip.QUERY_START | EndSkip.QUERY_STOP | EndSkip.REFERENCE_START
    PREFIX = EndSkip.QUERY_STOP
    SUFFIX = EndSkip.QUERY_START
    FRONT_NOT_INTERNAL = EndSkip.REFERENCE_START | EndSkip.QUERY_STOP
    BACK_NOT_INTERNAL = EndSkip.QUERY_START | EndSkip.REFERENCE_END
    ANYWHERE = EndSkip.SEMIGLOBAL


class SingleMatch:
    def __init__(self, astart, astop, rstart, rstop, score, errors, adapter, sequence):
        self.astart, self.astop = astart, astop
        self.rstart, self.rstop = rstart, rstop
        self.score, self.errors = score, errors
        self.adapter = adapter
        self.sequence = sequence
        self.length = astop - astart

    def __repr__(self):
        return (f"{type(self).__name__}(astart={self.astart}, astop={self.astop}, rstart={self.rstart}, "
                f"rstop={self.rstop}, score={self.score}, errors={self.errors})")

    def as_tuple(self):
        return (self.astart, self.astop, self.rstart, self.rstop, self.score, self.errors)


class RemoveBeforeMatch(SingleMatch):
    """5' kinds: the match and everything before it is removed."""

    def trimmed(self, read):
        return read[self.rstop:]

    def removed_sequence_length(self):
        return self.rstop


class RemoveAfterMatch(SingleMatch):
    """3' kinds: the match and everything after it is removed."""

    def trimmed(self, read):
        return read[: self.rstart]

    def removed_sequence_length(self):
        return len(self.sequence) - self.rstart


class SingleAdapter:
    where = None
    match_class = None
    description = "adapter"

    def __init__(self, sequence, max_errors=0.1, min_overlap=3, read_wildcards=False,
                 adapter_wildcards=True, name=None, indels=True):
        self.name = name
        self._debug = False
        self.sequence = sequence.upper().replace("U", "T")
        if not self.sequence:
            raise ValueError("Adapter sequence is empty")
        if max_errors >= 1:
            max_errors /= len(self.sequence)
        self.max_error_rate = max_errors
        self.min_overlap = min(min_overlap, len(self.sequence))
        self.read_wildcards = read_wildcards
        self.indels = indels
        # wildcards are switched off when the adapter has none
        self.adapter_wildcards = adapter_wildcards and not set(self.sequence) <= set("ACGT")
        if self.adapter_wildcards or self.read_wildcards or not self.indels:
            raise NotImplementedError("not on the cutseq path")
        self.aligner = self._aligner()

    def _make_aligner(self, sequence, flags):
        return Aligner(sequence, self.max_error_rate, flags=flags, wildcard_ref=False,
                       wildcard_query=False, indel_cost=1, min_overlap=self.min_overlap)

    def _aligner(self):
        return self._make_aligner(self.sequence, int(self.where))

    def __repr__(self):
        return (f"<{type(self).__name__}(name={self.name!r}, sequence={self.sequence!r}, "
                f"max_error_rate={self.max_error_rate}, min_overlap={self.min_overlap}, "
                f"read_wildcards={self.read_wildcards}, adapter_wildcards={self.adapter_wildcards}, "
                f"indels={self.indels})>")

    def __len__(self):
        return len(self.sequence)

    def match_to(self, sequence):
        alignment = self.aligner.locate(sequence.upper())
        if alignment is None:
            return None
        return self.match_class(*alignment, self, sequence)


class FrontAdapter(SingleAdapter):
    where = Where.FRONT
    match_class = RemoveBeforeMatch
    description = "regular 5'"

    def __init__(self, *args, **kwargs):
        self._force_anywhere = kwargs.pop("force_anywhere", False)
        super().__init__(*args, **kwargs)

    def _aligner(self):
        return self._make_aligner(self.sequence, int(Where.ANYWHERE if self._force_anywhere else self.where))


class RightmostFrontAdapter(FrontAdapter):
    """5' adapter preferring the rightmost occurrence: the reversed adapter is searched
    as a 3' (BACK) adapter in the reversed read and the coordinates are mapped back."""

    description = "rightmost 5'"

    def _aligner(self):
        return self._make_aligner(self.sequence[::-1], int(Where.ANYWHERE if self._force_anywhere else Where.BACK))

    def match_to(self, sequence):
        alignment = self.aligner.locate(sequence.upper()[::-1])
        if alignment is None:
            return None
        ref_start, ref_end, query_start, query_end, score, errors = alignment
        m, n = len(self.sequence), len(sequence)
        return RemoveBeforeMatch(m - ref_end, m - ref_start, n - query_end, n - query_start,
                                 score, errors, self, sequence)


class BackAdapter(SingleAdapter):
    where = Where.BACK
    match_class = RemoveAfterMatch
    description = "regular 3'"

    def __init__(self, *args, **kwargs):
        self._force_anywhere = kwargs.pop("force_anywhere", False)
        super().__init__(*args, **kwargs)

    def _aligner(self):
        return self._make_aligner(self.sequence, int(Where.ANYWHERE if self._force_anywhere else self.where))


class NonInternalFrontAdapter(FrontAdapter):
    where = Where.FRONT_NOT_INTERNAL
    description = "non-internal 5'"

    def _aligner(self):
        return self._make_aligner(self.sequence, int(Where.FRONT_NOT_INTERNAL))


class NonInternalBackAdapter(BackAdapter):
    where = Where.BACK_NOT_INTERNAL
    description = "non-internal 3'"

    def _aligner(self):
        return self._make_aligner(self.sequence, int(Where.BACK_NOT_INTERNAL))


class PrefixAdapter(NonInternalFrontAdapter):
    """Anchored 5' adapter; with indels the regular aligner is used with PREFIX flags and
    the minimum overlap is the full adapter length."""

    where = Where.PREFIX
    description = "anchored 5'"

    def __init__(self, sequence, *args, **kwargs):
        kwargs["min_overlap"] = len(sequence)
        super().__init__(sequence, *args, **kwargs)

    def _aligner(self):
        return self._make_aligner(self.sequence, int(Where.PREFIX))


class SuffixAdapter(NonInternalBackAdapter):
    where = Where.SUFFIX
    description = "anchored 3'"

    def __init__(self, sequence, *args, **kwargs):
        kwargs["min_overlap"] = len(sequence)
        super().__init__(sequence, *args, **kwargs)

    def _aligner(self):
        return self._make_aligner(self.sequence, int(Where.SUFFIX))


class MultipleAdapters:
    def __init__(self, adapters):
        self._adapters = list(adapters)

    def __iter__(self):
        return iter(self._adapters)

    def __len__(self):
        return len(self._adapters)

    def __getitem__(self, i):
        return self._adapters[i]

    def match_to(self, sequence):
        best = None
        for adapter in self._adapters:
            match = adapter.match_to(sequence)
            if match is None:
                continue
            if best is None or match.score > best.score or (match.score == best.score and match.errors < best.errors):
                best = match
        return best
