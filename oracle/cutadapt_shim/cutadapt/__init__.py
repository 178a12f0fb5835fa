"""TEST INFRASTRUCTURE ONLY - a pure-Python restatement of the parts of cutadapt 5.x
that the reference cutseq imports (reference cutseq/run.py:16-46).

cutadapt itself is an un-vendored dependency of the reference (pyproject.toml:17,
``cutadapt~=5.0``) and cannot be installed in the build image.  This package restates
its published algorithms from the upstream sources (src/cutadapt/_align.pyx,
adapters.py, modifiers.py, qualtrim.pyx, predicates.py, steps.py, pipeline.py,
info.py) so that the UNMODIFIED reference ``cutseq/run.py`` can be imported and run
on top of it to generate golden vectors (scripts/make_golden.py).

PARITY UNPINNED: the reference ships no expected outputs or known-answer tests, and no
real cutadapt was reachable to check this restatement against.

Nothing under cutseq_b200/ may import this package.
"""

__version__ = "5.0+restated"
