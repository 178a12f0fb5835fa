"""Restatement of upstream src/cutadapt/predicates.py (the parts cutseq uses)."""


class Predicate:
    def test(self, read, info) -> bool:
        raise NotImplementedError

    @classmethod
    def descriptive_identifier(cls):
        return "".join(("_" + ch.lower() if ch.isupper() else ch) for ch in cls.__name__)[1:]


class TooShort(Predicate):
    def __init__(self, minimum_length: int):
        self.minimum_length = minimum_length

    def __repr__(self):
        return f"TooShort(minimum_length={self.minimum_length})"

    def descriptive_identifier(self):
        return "too_short"

    def test(self, read, info):
        return len(read) < self.minimum_length
