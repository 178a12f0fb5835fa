"""Restatement of the modifiers cutseq uses (upstream src/cutadapt/modifiers.py)."""

from .adapters import MultipleAdapters
from .info import ModificationInfo
from .qualtrim import quality_trim_index
from ._record import record_names_match


class SingleEndModifier:
    def __call__(self, read, info: ModificationInfo):
        raise NotImplementedError


class PairedEndModifier:
    def __call__(self, read1, read2, info1, info2):
        raise NotImplementedError


class PairedEndModifierWrapper(PairedEndModifier):
    """Applies one single-end modifier per mate (either may be None)."""

    paired = True

    def __init__(self, modifier1, modifier2):
        if modifier1 is None and modifier2 is None:
            raise ValueError("Not both modifiers may be None")
        self._modifier1, self._modifier2 = modifier1, modifier2

    def __repr__(self):
        return f"PairedEndModifierWrapper({self._modifier1!r}, {self._modifier2!r})"

    def __call__(self, read1, read2, info1, info2):
        if self._modifier1 is None:
            return read1, self._modifier2(read2, info2)
        if self._modifier2 is None:
            return self._modifier1(read1, info1), read2
        return self._modifier1(read1, info1), self._modifier2(read2, info2)


class AdapterCutter(SingleEndModifier):
    def __init__(self, adapters, times=1, action="trim", index=True):
        if action != "trim":
            raise NotImplementedError("cutseq only uses the default action")
        self.times = times
        self.action = action
        self.with_adapters = 0
        self.adapters = MultipleAdapters(adapters)
        self.adapter_statistics = {a: {"matches": 0} for a in adapters}

    def __repr__(self):
        return f"AdapterCutter(adapters={list(self.adapters)!r}, times={self.times}, action='{self.action}')"

    def match_and_trim(self, read):
        matches = []
        trimmed_read = read
        for _ in range(self.times):
            match = self.adapters.match_to(trimmed_read.sequence)
            if match is None:
                break
            matches.append(match)
            trimmed_read = match.trimmed(trimmed_read)
        return trimmed_read, matches

    def __call__(self, read, info: ModificationInfo):
        trimmed_read, matches = self.match_and_trim(read)
        if matches:
            self.with_adapters += 1
            for match in matches:
                self.adapter_statistics[match.adapter]["matches"] += 1
        info.matches.extend(matches)
        return trimmed_read


class UnconditionalCutter(SingleEndModifier):
    def __init__(self, length: int):
        self.length = length

    def __repr__(self):
        return f"UnconditionalCutter(length={self.length})"

    def __call__(self, read, info: ModificationInfo):
        if self.length > 0:
            info.cut_prefix = read.sequence[: self.length]
            return read[self.length:]
        elif self.length < 0:
            info.cut_suffix = read.sequence[self.length:]
            return read[: self.length]
        return read


class SuffixRemover(SingleEndModifier):
    def __init__(self, suffix: str):
        self.suffix = suffix

    def __repr__(self):
        return f"SuffixRemover('{self.suffix}')"

    def __call__(self, read, info: ModificationInfo):
        read = read[:]
        if read.name.endswith(self.suffix):
            read.name = read.name[: -len(self.suffix)]
        return read


class Renamer(SingleEndModifier):
    def __init__(self, template: str):
        self._template = template.replace(r"\t", "\t")

    def __repr__(self):
        return f"Renamer('{self._template}')"

    @staticmethod
    def parse_name(read_name: str):
        fields = read_name.split(maxsplit=1)
        if len(fields) == 2:
            return (fields[0], fields[1])
        return (read_name, "")

    def __call__(self, read, info: ModificationInfo):
        id_, comment = self.parse_name(read.name)
        read.name = self._template.format(
            header=read.name, id=id_, comment=comment,
            cut_prefix=info.cut_prefix if info.cut_prefix else "",
            cut_suffix=info.cut_suffix if info.cut_suffix else "",
        )
        return read


class _InfoView:
    def __init__(self, info):
        self.cut_prefix = info.cut_prefix if info.cut_prefix else ""
        self.cut_suffix = info.cut_suffix if info.cut_suffix else ""


class PairedEndRenamer(PairedEndModifier):
    def __init__(self, template: str):
        self._template = template.replace(r"\t", "\t")

    def __repr__(self):
        return f"PairedEndRenamer('{self._template}')"

    def __call__(self, read1, read2, info1, info2):
        id1, comment1 = Renamer.parse_name(read1.name)
        id2, comment2 = Renamer.parse_name(read2.name)
        if not record_names_match(read1.name, read2.name):
            raise ValueError(f"Input read IDs not identical: '{id1}' != '{id2}'")
        r1, r2 = _InfoView(info1), _InfoView(info2)
        name1 = self._template.format(header=read1.name, id=id1, comment=comment1,
                                      cut_prefix=r1.cut_prefix, cut_suffix=r1.cut_suffix, r1=r1, r2=r2)
        name2 = self._template.format(header=read2.name, id=id2, comment=comment2,
                                      cut_prefix=r2.cut_prefix, cut_suffix=r2.cut_suffix, r1=r1, r2=r2)
        read1.name, read2.name = name1, name2
        return read1, read2


class QualityTrimmer(SingleEndModifier):
    def __init__(self, cutoff_front: int, cutoff_back: int, base: int = 33):
        self.cutoff_front, self.cutoff_back, self.base = cutoff_front, cutoff_back, base
        self.trimmed_bases = 0

    def __repr__(self):
        return f"QualityTrimmer(cutoff_front={self.cutoff_front}, cutoff_back={self.cutoff_back}, base={self.base})"

    def __call__(self, read, info: ModificationInfo):
        start, stop = quality_trim_index(read.qualities, self.cutoff_front, self.cutoff_back, self.base)
        self.trimmed_bases += len(read) - (stop - start)
        return read[start:stop]
