"""Stand-in for upstream src/cutadapt/report.py: the counters behind ``minimal_report``.
``_collect_modifier`` keeps upstream's one-AdapterCutter-per-mate assertion, which the
reference swallows with a monkey-patch (run.py:58-73)."""

from .modifiers import AdapterCutter, PairedEndModifierWrapper, QualityTrimmer
from .steps import PairedEndFilter, PairedEndSink, SingleEndFilter, SingleEndSink


class Statistics:
    def __init__(self):
        self.paired = None
        self.n = 0
        self.total_bp = [0, 0]
        self.with_adapters = [None, None]
        self.quality_trimmed_bp = [None, None]
        self.filtered = {}
        self.written = 0
        self.written_bp = [0, 0]
        self.adapter_stats = [[], []]

    @property
    def total(self):
        return sum(self.total_bp)

    @property
    def total_written_bp(self):
        return self.written_bp

    def collect(self, n, total_bp1, total_bp2, modifiers, steps):
        self.n = n
        self.total_bp[0] = total_bp1
        self.paired = total_bp2 is not None
        if total_bp2 is not None:
            self.total_bp[1] = total_bp2
        for modifier in modifiers:
            self._collect_modifier(modifier)
        for step in steps:
            self._collect_step(step)

    def _collect_step(self, step):
        if isinstance(step, (SingleEndSink, PairedEndSink)):
            self.written += step.written_reads
            self.written_bp[0] += step.written_bp[0]
            self.written_bp[1] += step.written_bp[1]
        elif isinstance(step, (SingleEndFilter, PairedEndFilter)):
            name = step.descriptive_identifier()
            self.filtered[name] = self.filtered.get(name, 0) + step.filtered

    def _collect_modifier(self, m):
        if isinstance(m, PairedEndModifierWrapper):
            for i, sub in enumerate((m._modifier1, m._modifier2)):
                self._collect_single(sub, i)
        else:
            self._collect_single(m, 0)

    def _collect_single(self, m, i):
        if isinstance(m, QualityTrimmer):
            self.quality_trimmed_bp[i] = (self.quality_trimmed_bp[i] or 0) + m.trimmed_bases
        elif isinstance(m, AdapterCutter):
            assert self.with_adapters[i] is None
            self.with_adapters[i] = m.with_adapters
            self.adapter_stats[i] = list(m.adapter_statistics.values())

    def as_json(self, *args, **kwargs):
        return {
            "read_counts": {"input": self.n, "output": self.written,
                            "filtered": dict(self.filtered)},
            "basepair_counts": {"input": self.total, "output": sum(self.written_bp)},
        }


def minimal_report(stats, time=None, gc_content=None) -> str:
    fields = [
        "OK", stats.n, stats.total,
        stats.filtered.get("too_short", 0) or 0,
        stats.filtered.get("too_long", 0) or 0,
        stats.filtered.get("too_many_n", 0) or 0,
        stats.written,
        stats.with_adapters[0] if stats.with_adapters[0] is not None else 0,
        stats.quality_trimmed_bp[0] if stats.quality_trimmed_bp[0] is not None else 0,
        stats.total_written_bp[0],
    ]
    header = ["status", "in_reads", "in_bp", "too_short", "too_long", "too_many_n",
              "out_reads", "w/adapters", "qualtrim_bp", "out_bp"]
    if stats.paired:
        fields += [
            stats.with_adapters[1] if stats.with_adapters[1] is not None else 0,
            stats.quality_trimmed_bp[1] if stats.quality_trimmed_bp[1] is not None else 0,
            stats.total_written_bp[1],
        ]
        header += ["w/adapters2", "qualtrim_bp2", "out2_bp"]
    return "\t".join(header) + "\n" + "\t".join(str(x) for x in fields)
