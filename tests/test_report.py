"""Reports of the run (reference run.py:222-302 ``json_report``, run.py:489 / 810 ``minimal_report``) from a counters
structure - no GPU needed: the counters are filled by hand."""
import json

from cutseq_b200 import _abi as A
from cutseq_b200 import program, run
from cutseq_b200.common import BarcodeConfig


def _setup(argv, paired=True):
    args = run.build_parser().parse_args(argv + (["a.fq", "b.fq"] if paired else ["a.fq"]))
    bc = BarcodeConfig(run.resolve_scheme(args))
    settings = run.settings_from_args(args)
    prog = program.compile_paired(bc, settings) if paired else program.compile_single(bc, settings)
    c = A.csq_counters()
    c.n, c.written, c.too_short, c.untrimmed = 1000, 940, 50, 10
    c.total_bp[0], c.total_bp[1] = 150000, 149000 if paired else 0
    c.written_bp[0], c.written_bp[1] = 120000, 118000 if paired else 0
    c.quality_trimmed_bp[0], c.quality_trimmed_bp[1] = 700, 900 if paired else 0
    for m, ops in enumerate((prog.ops_r1, prog.ops_r2)):
        for i in range(len(ops)):
            c.with_adapters[m][i] = 100 * (m + 1) + i
    return bc, prog, c


def test_error_lengths_follow_cutadapt_error_ranges():
    assert run._error_lengths(20, 0.2) == [4, 9, 14, 19]
    assert run._error_lengths(20, 0.1) == [9, 19]
    assert run._error_lengths(6, 0.2) == [4]
    assert run._error_lengths(3, 0.2) == []


def test_json_report_has_cutadapt_layout(tmp_path):
    bc, prog, c = _setup(["-A", "TAKARAV3", "--trim-polyA"])
    f = str(tmp_path / "r.json")
    run.json_report(f, c, prog, bc, "a.fq", "b.fq", "o1", "o2", "s1", "s2", None, None)
    d = json.load(open(f))
    assert d["tag"] == "Cutadapt report" and d["input"] == {"path1": "a.fq", "path2": "b.fq", "paired": True}
    assert d["barcode"]["p5"] == "ACACGACGCTCTTCCGATCT" and d["barcode"]["strand"] == "-"
    rc, bp = d["read_counts"], d["basepair_counts"]
    assert (rc["input"], rc["output"], rc["filtered"]["too_short"], rc["filtered"]["is_untrimmed_any"]) == (1000, 940, 50, 10)
    assert rc["filtered"]["too_long"] is None and rc["reverse_complemented"] is None
    assert bp["input"] == 299000 and bp["output"] == 238000 and bp["quality_trimmed"] == 1600
    # only the FIRST AdapterCutter of each mate reaches cutadapt's statistics (the monkey-patch quirk, run.py:58-73)
    first = [next(i for i, op in enumerate(ops) if op.kind == A.OP_ALIGN) for ops in (prog.ops_r1, prog.ops_r2)]
    for m, key in enumerate(("adapters_read1", "adapters_read2")):
        (entry,) = d[key]
        assert entry["name"] == str(m + 1) and entry["linked"] is False and entry["three_prime_end"] is None
        end = entry["five_prime_end"]
        assert end["type"] == "rightmost_five_prime" and end["error_rate"] == 0.2 and end["error_lengths"] == [4, 9, 14, 19]
        assert end["matches"] == entry["total_matches"] == 100 * (m + 1) + first[m] and end["trimmed_lengths"] == []
    assert rc["read1_with_adapter"] == 100 + first[0] and rc["read2_with_adapter"] == 200 + first[1]
    assert d["adapters_read1"][0]["five_prime_end"]["sequence"] == "ACACGACGCTCTTCCGATCT"
    assert d["adapters_read2"][0]["five_prime_end"]["sequence"] == "GACGTGTGCTCTTCCGATCT"  # p7 reverse complement


def test_json_report_single_end(tmp_path):
    bc, prog, c = _setup(["-A", "SMALLRNA"], paired=False)
    f = str(tmp_path / "r.json")
    run.json_report(f, c, prog, bc, "a.fq", None, "o1", None, "s1", None, None, None)
    d = json.load(open(f))
    assert d["input"]["paired"] is False and d["adapters_read2"] is None and d["read_counts"]["read2_with_adapter"] is None
    assert d["basepair_counts"]["input"] == 150000 and d["basepair_counts"]["input_read2"] is None
    assert d["basepair_counts"]["quality_trimmed_read2"] is None


def test_json_report_three_prime_end_has_adjacent_bases(tmp_path):
    """A first cutter that is a 3' adapter (every scheme of run.py starts with the 5' p5 adapter, so only programs built
    through the C ABI get here): cutadapt keeps the base in front of every match (EndStatistics.adjacent_bases) and
    names a dominant one (> 80 % of >= 20 matches)."""
    bc, prog, c = _setup(["-A", "SMALLRNA"], paired=False)
    first = next(op for op in prog.ops_r1 if op.kind == A.OP_ALIGN)
    first.adapter_kind = A.AD_BACK
    for slot, v in enumerate((90, 3, 2, 1, 2, 1)):  # A, C, G, T, none, other
        c.adjacent_bases[0][slot] = v
    f = str(tmp_path / "r.json")
    run.json_report(f, c, prog, bc, "a.fq", None, "o1", None, "s1", None, None, None)
    end = json.load(open(f))["adapters_read1"][0]["three_prime_end"]
    assert end["adjacent_bases"] == {"A": 90, "C": 3, "G": 2, "T": 1, "": 3} and end["dominant_adjacent_base"] == "A"
    c.adjacent_bases[0][0] = 5
    run.json_report(f, c, prog, bc, "a.fq", None, "o1", None, "s1", None, None, None)
    end = json.load(open(f))["adapters_read1"][0]["three_prime_end"]
    assert end["dominant_adjacent_base"] is None  # 14 matches: too few to call


def test_minimal_report_uses_the_first_cutter_only():
    bc, prog, c = _setup(["-A", "TAKARAV3"])
    header, values = run.minimal_report_text(c, prog).split("\n")
    assert header.split("\t")[:4] == ["status", "in_reads", "in_bp", "too_short"] and header.endswith("out2_bp")
    v = values.split("\t")
    first = [next(i for i, op in enumerate(ops) if op.kind == A.OP_ALIGN) for ops in (prog.ops_r1, prog.ops_r2)]
    assert v[0] == "OK" and int(v[1]) == 1000 and int(v[2]) == 299000 and int(v[3]) == 50 and int(v[6]) == 940
    assert int(v[7]) == 100 + first[0] and int(v[10]) == 200 + first[1]
