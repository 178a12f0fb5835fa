"""Stand-in for upstream src/cutadapt/runners.py: a serial runner only."""

from ._record import read_fastq, record_names_match
from .report import Statistics


class _Format:
    def has_qualities(self):
        return True


class SerialPipelineRunner:
    def __init__(self, inpaths):
        self._inpaths = inpaths

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def input_file_format(self):
        return _Format()

    def run(self, pipeline, progress, outfiles):
        paths = self._inpaths.paths
        if len(paths) == 2:
            def pairs():
                it1, it2 = read_fastq(paths[0]), read_fastq(paths[1])
                while True:
                    r1, r2 = next(it1, None), next(it2, None)
                    if r1 is None and r2 is None:
                        return
                    if r1 is None or r2 is None:
                        raise ValueError("paired input files have different numbers of records")
                    if not record_names_match(r1.name, r2.name):  # dnaio TwoFilePairedEndReader: r1.is_mate(r2)
                        raise ValueError(f"Records are improperly paired. Read name '{r1.name}' in file 1 does not match "
                                         f"'{r2.name}' in file 2.")
                    yield r1, r2
            n, bp1, bp2 = pipeline.process_reads(pairs(), progress)
        else:
            n, bp1, bp2 = pipeline.process_reads(read_fastq(paths[0]), progress)
        stats = Statistics()
        stats.collect(n, bp1, bp2, pipeline._modifiers, pipeline._steps)
        return stats


def make_runner(inpaths, cores=1, buffer_size=None):
    return SerialPipelineRunner(inpaths)
