"""Op-program compiler: (BarcodeConfig, CutadaptConfig) -> the per-mate op lists and
filters that the GPU chain executes.

This is the host-side mirror of the modifier/step assembly in the reference:
single-end ``run.py:326-426`` (modifiers) + ``446-471`` (steps), paired-end
``run.py:533-731`` + ``763-792``.  Every op is one entry of run.py's ``modifiers`` list,
in the same order, with the same parameters; SURVEY.md Table 8.1 is the row-by-row map.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

from . import _abi as A

MAX_ERRORS = 0.2  # run.py:326, 533
POLYA_MAX_ERRORS = 0.15  # run.py:389, 674
POLYA_LENGTH = 100  # run.py:390, 675
ID_INLINE5 = 0  # adapter_id bits used by the IsUntrimmedAny filter
ID_INLINE3 = 1


@dataclass
class Op:
    kind: int
    adapter_kind: int = 0
    adapter: str = ""
    min_overlap: int = 3  # cutadapt's default
    max_error_rate: float = 0.0
    adapter_id: int = -1
    length: int = 0
    force_trim_min_length: int = 0
    suffix: str = ""
    rename_parts: int = 0
    cutoff_front: int = 0
    cutoff_back: int = 0
    quality_base: int = 33

    def to_c(self) -> A.csq_op:
        c = A.csq_op()
        c.kind = self.kind
        c.adapter_kind = self.adapter_kind
        if self.kind == A.OP_ALIGN:
            if not 1 <= len(self.adapter) <= A.CSQ_MAX_ADAPTER:
                raise ValueError(f"adapter length {len(self.adapter)} outside 1..{A.CSQ_MAX_ADAPTER}")
        c.adapter_len = len(self.adapter)
        c.adapter = self.adapter.encode("ascii")
        c.min_overlap = self.min_overlap
        c.adapter_id = self.adapter_id
        c.max_error_rate = self.max_error_rate
        c.length = self.length
        c.force_trim_min_length = self.force_trim_min_length
        c.suffix_len = len(self.suffix)
        c.suffix = self.suffix.encode("ascii")
        c.rename_parts = self.rename_parts
        c.cutoff_front = self.cutoff_front
        c.cutoff_back = self.cutoff_back
        c.quality_base = self.quality_base
        return c

    def __str__(self) -> str:  # --dry-run listing
        k = self.kind
        if k == A.OP_STRIP_SUFFIX:
            return f"SuffixRemover('{self.suffix}')"
        if k == A.OP_ALIGN:
            return (
                f"AdapterCutter({A.AD_NAMES[self.adapter_kind]}(sequence='{self.adapter}', "
                f"max_error_rate={self.max_error_rate}, min_overlap={self.min_overlap}), times=1, action='trim')"
            )
        if k == A.OP_CUT:
            return f"UnconditionalCutter(length={self.length})"
        if k == A.OP_COND_CUT:
            return f"ConditionalCutter(length={self.length}, force_trim_min_length={self.force_trim_min_length})"
        if k == A.OP_RENAME:
            return f"Renamer('{rename_template(self.rename_parts)}')"
        if k == A.OP_QTRIM:
            return f"QualityTrimmer(cutoff_front={self.cutoff_front}, cutoff_back={self.cutoff_back}, base={self.quality_base})"
        if k == A.OP_REVCOMP:
            return "ReverseComplementConverter()"
        return f"Op({k})"


def rename_template(parts: int) -> str:
    if parts == 0:
        return "{id}"
    t = "{id}_"
    if parts & A.REN_OWN_PREFIX:
        t += "{cut_prefix}"
    if parts & A.REN_OWN_SUFFIX:
        t += "{cut_suffix}"
    if parts & A.REN_R1_PREFIX:
        t += "{r1.cut_prefix}"
    if parts & A.REN_R2_PREFIX:
        t += "{r2.cut_prefix}"
    return t


@dataclass
class Filters:
    min_length: int = 20
    untrimmed_enabled: bool = False
    required_r1: int = 0
    required_r2: int = 0

    def to_c(self) -> A.csq_filters:
        return A.csq_filters(self.min_length, int(self.untrimmed_enabled), self.required_r1, self.required_r2)


@dataclass
class Program:
    paired: bool
    ops_r1: List[Op]
    ops_r2: List[Op] = field(default_factory=list)
    filters: Filters = field(default_factory=Filters)
    swap_sink: bool = False  # paired --auto-rc on a '-' strand library (run.py:785-792)
    notes: List[str] = field(default_factory=list)  # log lines the reference emits while assembling

    def c_ops(self, mate: int):
        ops = self.ops_r1 if mate == 0 else self.ops_r2
        arr = (A.csq_op * max(1, len(ops)))()
        for i, op in enumerate(ops):
            arr[i] = op.to_c()
        return arr, len(ops)

    def describe(self) -> List[str]:
        lines = []
        if self.paired:
            for i, (a, b) in enumerate(zip(self.ops_r1, self.ops_r2), 1):
                lines.append(f"Step {i}: ({a}, {b})" if str(a) != str(b) or a.kind != A.OP_RENAME else f"Step {i}: Paired{a}")
        else:
            for i, a in enumerate(self.ops_r1, 1):
                lines.append(f"Step {i}: {a}")
        return lines


def _align(kind: int, seq: str, rate: float, min_overlap: int = 3, adapter_id: int = -1) -> Op:
    return Op(A.OP_ALIGN, adapter_kind=kind, adapter=seq, min_overlap=min_overlap, max_error_rate=rate, adapter_id=adapter_id)


def _cut(length: int) -> Op:
    return Op(A.OP_CUT, length=length)


def _maybe_cond_cut(length: int, settings) -> Op:
    # run.py:622-627 etc.: ConditionalCutter unless --no-conditional-cutter
    if settings.conditional_cutter:
        return Op(A.OP_COND_CUT, length=length, force_trim_min_length=settings.force_trim_min_length)
    return _cut(length)


def compile_single(barcode, settings, untrimmed1: Optional[str] = None) -> Program:
    """run.py:326-426 (modifiers) and 446-471 (steps)."""
    ops: List[Op] = []
    notes: List[str] = []
    back_kind = A.AD_BACK_ANYWHERE if settings.force_anywhere else A.AD_BACK
    # step 1
    ops += [Op(A.OP_STRIP_SUFFIX, suffix=".1"), Op(A.OP_STRIP_SUFFIX, suffix="/1")]
    # step 2, 3
    ops.append(_align(A.AD_RIGHTMOST_FRONT, barcode.p5.fw, MAX_ERRORS, 10))
    ops.append(_align(back_kind, barcode.p7.fw, MAX_ERRORS, 3))
    # step 4
    required = 0
    if barcode.inline5.len > 0:
        ops.append(_align(A.AD_PREFIX, barcode.inline5.fw, MAX_ERRORS, adapter_id=ID_INLINE5))
        required |= 1 << ID_INLINE5
    if barcode.inline3.len > 0:
        ops.append(_align(A.AD_SUFFIX, barcode.inline3.fw, MAX_ERRORS, adapter_id=ID_INLINE3))
        required |= 1 << ID_INLINE3
    # step 5
    if barcode.umi5.len > 0:
        ops.append(_cut(barcode.umi5.len))
    if barcode.umi3.len > 0:
        ops.append(_cut(-barcode.umi3.len))
    with_umi = barcode.umi5.len + barcode.umi3.len > 0
    ops.append(Op(A.OP_RENAME, rename_parts=(A.REN_OWN_PREFIX | A.REN_OWN_SUFFIX) if with_umi else 0))
    # step 6
    if barcode.mask5.len > 0:
        ops.append(_cut(barcode.mask5.len))
    if barcode.mask3.len > 0:
        ops.append(_cut(-barcode.mask3.len))
    # step 7
    if settings.trim_polyA:
        fwd = _align(A.AD_NI_BACK, "A" * POLYA_LENGTH, POLYA_MAX_ERRORS)
        rev = _align(A.AD_NI_FRONT, "T" * POLYA_LENGTH, POLYA_MAX_ERRORS)
        if settings.trim_polyA_wo_direction:
            ops += [fwd, rev]
        elif barcode.strand == "+":
            ops.append(fwd)
        elif barcode.strand == "-":
            ops.append(rev)
        else:
            notes.append("INFO:No strand information provided, skip polyA trimming.")
    # step 8
    ops.append(Op(A.OP_QTRIM, cutoff_front=0, cutoff_back=settings.min_quality))
    # step 9
    if settings.auto_rc:
        if barcode.strand == "-":
            ops.append(Op(A.OP_REVCOMP))
        else:
            notes.append("WARNING:Library is not (-) strand, but --auto-rc is enabled. Ignored.")
    enabled = (barcode.inline5.len + barcode.inline3.len > 0 and settings.ensure_inline_barcode) or (
        untrimmed1 is not None
    )
    flt = Filters(settings.min_length, bool(enabled), required, 0)
    return Program(False, ops, [], flt, False, notes)


def compile_paired(barcode, settings, untrimmed1: Optional[str] = None, untrimmed2: Optional[str] = None) -> Program:
    """run.py:533-731 (modifiers) and 763-792 (steps)."""
    r1: List[Op] = []
    r2: List[Op] = []
    notes: List[str] = []
    back_kind = A.AD_BACK_ANYWHERE if settings.force_anywhere else A.AD_BACK

    def both(a: Op, b: Op):
        r1.append(a)
        r2.append(b)

    # step 1
    both(Op(A.OP_STRIP_SUFFIX, suffix=".1"), Op(A.OP_STRIP_SUFFIX, suffix=".2"))
    both(Op(A.OP_STRIP_SUFFIX, suffix="/1"), Op(A.OP_STRIP_SUFFIX, suffix="/2"))
    # step 2
    both(_align(A.AD_RIGHTMOST_FRONT, barcode.p5.fw, MAX_ERRORS, 10), _align(A.AD_RIGHTMOST_FRONT, barcode.p7.rc, MAX_ERRORS, 10))
    # step 3
    both(_align(back_kind, barcode.p7.fw, MAX_ERRORS, 3), _align(back_kind, barcode.p5.rc, MAX_ERRORS, 3))
    # step 4
    req1 = req2 = 0
    if barcode.inline5.len > 0:
        both(_align(A.AD_PREFIX, barcode.inline5.fw, MAX_ERRORS, adapter_id=ID_INLINE5), _cut(-barcode.inline5.len))
        req1 |= 1 << ID_INLINE5
    if barcode.inline3.len > 0:
        both(_cut(-barcode.inline3.len), _align(A.AD_PREFIX, barcode.inline3.rc, MAX_ERRORS, adapter_id=ID_INLINE3))
        req2 |= 1 << ID_INLINE3
    # step 5
    if barcode.umi5.len > 0:
        both(_cut(barcode.umi5.len), _maybe_cond_cut(-barcode.umi5.len, settings))
    if barcode.umi3.len > 0:
        both(_maybe_cond_cut(-barcode.umi3.len, settings), _cut(barcode.umi3.len))
    with_umi = barcode.umi5.len + barcode.umi3.len > 0
    parts = (A.REN_R1_PREFIX | A.REN_R2_PREFIX) if with_umi else 0
    both(Op(A.OP_RENAME, rename_parts=parts), Op(A.OP_RENAME, rename_parts=parts))
    # step 6
    if barcode.mask5.len > 0:
        both(_cut(barcode.mask5.len), _maybe_cond_cut(-barcode.mask5.len, settings))
    if barcode.mask3.len > 0:
        both(_maybe_cond_cut(-barcode.mask3.len, settings), _cut(barcode.mask3.len))
    # step 7
    if settings.trim_polyA:
        def polya_back():
            return _align(A.AD_NI_BACK, "A" * POLYA_LENGTH, POLYA_MAX_ERRORS)

        def polyt_front():
            return _align(A.AD_NI_FRONT, "T" * POLYA_LENGTH, POLYA_MAX_ERRORS)

        if settings.trim_polyA_wo_direction:
            both(polya_back(), polyt_front())
            both(polyt_front(), polya_back())
        elif barcode.strand == "+":
            both(polya_back(), polyt_front())
        elif barcode.strand == "-":
            both(polyt_front(), polya_back())
        else:
            notes.append("INFO:No strand information provided, skip polyA trimming.")
    # step 8
    both(
        Op(A.OP_QTRIM, cutoff_front=0, cutoff_back=settings.min_quality),
        Op(A.OP_QTRIM, cutoff_front=0, cutoff_back=settings.min_quality),
    )
    # step 9: no modifier; the sink swaps R1/R2 instead (run.py:727-731, 785-792)
    swap = False
    if settings.auto_rc:
        if barcode.strand != "-":
            notes.append("WARNING:Library is not (-) strand, but --auto-rc is enabled. Ignored.")
        else:
            swap = True
    enabled = (barcode.inline5.len + barcode.inline3.len > 0 and settings.ensure_inline_barcode) or (
        untrimmed1 is not None and untrimmed2 is not None
    )
    flt = Filters(settings.min_length, bool(enabled), req1, req2)
    return Program(True, r1, r2, flt, swap, notes)
