"""Stand-in for upstream src/cutadapt/utils.py ``Progress`` (cosmetic tty meter)."""


class Progress:
    def __init__(self, every=1):
        pass

    def update(self, increment, _final=False):
        pass

    def close(self):
        pass
