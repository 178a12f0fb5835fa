"""Restatement of upstream src/cutadapt/steps.py filters and sinks (as used by cutseq)."""


class SingleEndFilter:
    def __init__(self, predicate, writer):
        self.filtered = 0
        self.predicate, self.writer = predicate, writer

    def descriptive_identifier(self):
        return self.predicate.descriptive_identifier()

    def __call__(self, read, info):
        if self.predicate.test(read, info):
            self.filtered += 1
            if self.writer is not None:
                self.writer.write(read)
            return None
        return read


class PairedEndFilter:
    def __init__(self, predicate1, predicate2, writer, pair_filter_mode="any"):
        if pair_filter_mode not in ("any", "both", "first"):
            raise ValueError("pair_filter_mode must be 'any', 'both' or 'first'")
        self.filtered = 0
        self.predicate1, self.predicate2, self.writer = predicate1, predicate2, writer
        self._mode = pair_filter_mode

    def descriptive_identifier(self):
        p = self.predicate1 if self.predicate1 is not None else self.predicate2
        return p.descriptive_identifier()

    def _is_filtered(self, read1, read2, info1, info2):
        if self.predicate2 is None:
            return self.predicate1.test(read1, info1)
        if self.predicate1 is None:
            return self.predicate2.test(read2, info2)
        if self._mode == "any":
            return self.predicate1.test(read1, info1) or self.predicate2.test(read2, info2)
        if self._mode == "both":
            return self.predicate1.test(read1, info1) and self.predicate2.test(read2, info2)
        return self.predicate1.test(read1, info1)

    def __call__(self, read1, read2, info1, info2):
        if self._is_filtered(read1, read2, info1, info2):
            self.filtered += 1
            if self.writer is not None:
                self.writer.write(read1, read2)
            return None
        return (read1, read2)


class SingleEndSink:
    def __init__(self, writer):
        self.writer = writer
        self.written_reads = 0
        self.written_bp = [0, 0]

    def __call__(self, read, info):
        self.written_reads += 1
        self.written_bp[0] += len(read)
        self.writer.write(read)
        return None


class PairedEndSink:
    def __init__(self, writer):
        self.writer = writer
        self.written_reads = 0
        self.written_bp = [0, 0]

    def __call__(self, read1, read2, info1, info2):
        self.written_reads += 1
        self.written_bp[0] += len(read1)
        self.written_bp[1] += len(read2)
        self.writer.write(read1, read2)
        return None
