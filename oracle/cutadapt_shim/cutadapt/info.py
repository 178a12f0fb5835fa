"""Restatement of upstream src/cutadapt/info.py ``ModificationInfo``."""


class ModificationInfo:
    __slots__ = ["matches", "original_read", "cut_prefix", "cut_suffix", "is_rc"]

    def __init__(self, read):
        self.matches = []
        self.original_read = read
        self.cut_prefix = None
        self.cut_suffix = None
        self.is_rc = None
