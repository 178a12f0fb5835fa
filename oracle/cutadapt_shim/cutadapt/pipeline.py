"""Restatement of upstream src/cutadapt/pipeline.py per-read driver loops."""

from .info import ModificationInfo
from .modifiers import PairedEndModifier, PairedEndModifierWrapper


class SingleEndPipeline:
    paired = False

    def __init__(self, modifiers, steps):
        self._modifiers = list(modifiers)
        self._steps = list(steps)

    def process_reads(self, infiles, progress=None):
        n = total_bp = 0
        for read in infiles:
            n += 1
            total_bp += len(read)
            info = ModificationInfo(read)
            for modifier in self._modifiers:
                read = modifier(read, info)
            for step in self._steps:
                read = step(read, info)
                if read is None:
                    break
        return (n, total_bp, None)


class PairedEndPipeline:
    paired = True

    def __init__(self, modifiers, steps):
        self._modifiers = []
        for modifier in modifiers:
            if isinstance(modifier, tuple):
                self._modifiers.append(PairedEndModifierWrapper(*modifier))
            else:
                assert isinstance(modifier, PairedEndModifier)
                self._modifiers.append(modifier)
        self._steps = list(steps)

    def process_reads(self, infiles, progress=None):
        n = total1_bp = total2_bp = 0
        for read1, read2 in infiles:
            n += 1
            total1_bp += len(read1)
            total2_bp += len(read2)
            info1, info2 = ModificationInfo(read1), ModificationInfo(read2)
            for modifier in self._modifiers:
                read1, read2 = modifier(read1, read2, info1, info2)
            reads = (read1, read2)
            for step in self._steps:
                reads = step(reads[0], reads[1], info1, info2)
                if reads is None:
                    break
        return (n, total1_bp, total2_bp)
