"""Stand-ins for upstream src/cutadapt/files.py InputPaths / OutputFiles."""

from ._record import open_maybe_gz


class InputPaths:
    def __init__(self, *paths, interleaved=False):
        self.paths = paths
        self.interleaved = interleaved


class _SingleWriter:
    def __init__(self, f):
        self._f = f

    def write(self, read):
        self._f.write(read.fastq_bytes())


class _PairedWriter:
    def __init__(self, f1, f2):
        self._f1, self._f2 = f1, f2

    def write(self, read1, read2):
        self._f1.write(read1.fastq_bytes())
        self._f2.write(read2.fastq_bytes())


class OutputFiles:
    def __init__(self, *, proxied, qualities, interleaved):
        self._files = []

    def open_record_writer(self, *paths, interleaved=False, force_fasta=False):
        files = [open_maybe_gz(p, "wb") for p in paths]
        self._files.extend(files)
        if len(files) == 1:
            return _SingleWriter(files[0])
        return _PairedWriter(*files)

    def close(self):
        for f in self._files:
            f.close()
        self._files = []
