#!/usr/bin/env python
"""Run the UNMODIFIED reference cutseq (``/root/reference/cutseq/run.py``) on top of the
restated cutadapt under ``oracle/cutadapt_shim`` (build container only: the reference tree
does not exist on the GPU box).

    python scripts/run_reference.py -A TAKARAV3 -O /tmp/out R1.fq.gz R2.fq.gz

Every line of cutseq's own logic (flag handling, operation order, ConditionalCutter,
IsUntrimmedAny, sink swapping, output naming) is the reference's; only the cutadapt
internals are the restatement (PARITY UNPINNED, see oracle/README.md).
"""

import importlib.metadata
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("CUTSEQ_REFERENCE", "/root/reference")


def load_reference_main():
    sys.path.insert(0, os.path.join(HERE, "..", "oracle", "cutadapt_shim"))
    sys.path.insert(0, REFERENCE)
    real_version = importlib.metadata.version

    def version(name):  # the reference is not pip-installed here (run.py:190)
        if name in ("cutseq", "cutseq.run"):
            return "0.0.68"
        return real_version(name)

    importlib.metadata.version = version
    import cutseq.run as ref_run

    assert os.path.abspath(ref_run.__file__).startswith(os.path.abspath(REFERENCE))
    return ref_run


def main(argv=None):
    ref_run = load_reference_main()
    if argv is not None:
        sys.argv = ["cutseq"] + list(argv)
    ref_run.main()


if __name__ == "__main__":
    main()
