"""cutseq_b200: B200-native implementation of cutseq's per-read trimming pipeline.

Host Python keeps cutseq's surface (CLI, ``-A`` table, ``-a`` scheme grammar, fixed
per-read operation order); the per-read work runs in hand-written sm_100a CUDA
kernels behind the C ABI declared in ``include/cutseq_b200.h``.
"""

__version__ = "0.1.0"
