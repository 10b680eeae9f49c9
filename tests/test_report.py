"""bathsearch's main report, alignments included: Search.report() / bathhost_search_format_report against the
outputs the reference ships in tutorial/.

  * AMP_N-fs.out (bathsearch --fs: one hit with 6 frameshifts and a stop codon) and PTH2.out (default pipeline: 4 hits on both strands,
    model with a CS line): everything between the "Query:" block and the pipeline statistics -- the "Scores for complete hits" table,
    the per-hit table line and every alignment block -- byte for byte;
  * AMP_N-frameline.out (--fs --frameline), AMP_N.out and MET-ct4.out (2 queries, codon table 4, 6 hits) were written by an earlier program version whose per-hit table has two
    more columns: their alignment blocks (model / match / translation / codon / PP lines with coordinates) byte for byte.
The CPU tests put the oracle's stage calls behind the host pipeline; the gpu tests run the product path."""
import re

import pytest

import common


def run(hostapi, hmm, fasta, index=0, backend=None, gpu_ctx=None, whole=False, **opt):
    model = hostapi.QueryModel(common.golden(hmm), index)
    search = hostapi.Search(model, gpu_ctx=gpu_ctx, backend=backend, **opt)
    for name, seq in hostapi.read_fasta(common.golden(fasta)):
        search.add_sequence(name, hostapi.digitize_dna(seq))
    search.finish()
    text = search.output() if whole else search.report()
    search.close()
    return text


def hit_section(path):
    txt = open(path).read()
    return txt[txt.index("Scores for complete hits:"):txt.index("Internal pipeline statistics summary:")]


def query_section(path):
    """from "Query:" to the "Total number of hits:" line: everything but the banner, the option echo and the timings"""
    txt = open(path).read()
    end = txt.index("\n", txt.index("Total number of hits:")) + 1
    return txt[txt.index("Query:"):end]


def alignment_blocks(txt):
    return [m.group(1) for m in re.finditer(r"  Alignment:\n(.*?)\n\n(?=>>|\n|Internal)", txt, re.S)]


def check_all(hostapi, **where):
    assert run(hostapi, "AMP_N.bhmm", "target-AMP_N.fa", **where) == hit_section(common.golden("AMP_N-fs.out"))
    assert run(hostapi, "PTH2.bhmm", "target-PTH2.fa", std_only=1, **where) == hit_section(common.golden("PTH2.out"))
    assert run(hostapi, "AMP_N.bhmm", "target-AMP_N.fa", whole=True, **where) == query_section(common.golden("AMP_N-fs.out"))
    assert run(hostapi, "PTH2.bhmm", "target-PTH2.fa", whole=True, std_only=1, **where) == query_section(common.golden("PTH2.out"))
    got = alignment_blocks(run(hostapi, "AMP_N.bhmm", "target-AMP_N.fa", std_only=1, **where))
    assert len(got) == 1 and got == alignment_blocks(open(common.golden("AMP_N.out")).read())
    got = alignment_blocks(run(hostapi, "AMP_N.bhmm", "target-AMP_N.fa", show_frameline=1, **where))       # bathsearch --fs --frameline
    assert len(got) == 1 and " FRAME\n" in got[0] and got == alignment_blocks(open(common.golden("AMP_N-frameline.out")).read())
    sections = open(common.golden("MET-ct4.out")).read().split("Query:")[1:]
    for q in range(2):
        got = alignment_blocks(run(hostapi, "MET-ct4.bhmm", "target-MET.fa", index=q, std_only=1, **where))
        assert len(got) == 3 and got == alignment_blocks(sections[q])


def test_report_matches_shipped_outputs_cpu_backend(oracle):
    from bath_b200 import hostapi
    be, keep = oracle.cpu_backend(4)
    check_all(hostapi, backend=be)
    del keep


def test_fstblout_lists_the_frameshifts_of_the_cigar(oracle):
    """--fstblout (no shipped example): the AMP_N hit's table must list, in order, exactly the frameshifts of its CIGAR string
    44M1F39M1B114M9I25M2B19M1B44M1B4M6I30M2B67M (F = 1-nt insertion, nB = n-nt deletion) at increasing target positions."""
    from bath_b200 import hostapi
    be, keep = oracle.cpu_backend(2)
    model = hostapi.QueryModel(common.golden("AMP_N.bhmm"))
    search = hostapi.Search(model, backend=be)
    for name, seq in hostapi.read_fasta(common.golden("target-AMP_N.fa")):
        search.add_sequence(name, hostapi.digitize_dna(seq))
    hit = search.finish()[0]
    lines = search.fstblout().splitlines()
    assert lines[0].startswith("# target name") and lines[1].startswith("#---")
    rows = [l.split() for l in lines[2:]]
    assert all(r[0] == "seq1" and r[2] == "AMP_N" and (int(r[5]), int(r[6])) == (1, 402) for r in rows)
    events = [(r[7], int(r[8])) for r in rows if r[7] != "S"]
    want, pos, starts = [], 1, []
    for n, op in re.findall(r"(\d+)([MIDFB])", hit["cigar"]):
        n = int(n)
        if op == "F":
            want.append(("I", n))
        elif op == "B":
            want.append(("D", n))
    # the hit's one stop codon sits in an insert state (model '.', translation '*' in AMP_N-fs.out): the reference lists stop codons of
    # match states only (src/p7_tophits.c:1517), so no 'S' line here
    assert events == want and sum(1 for r in rows if r[7] == "S") == 0 and hit["stops"] == 1
    assert int(rows[0][9]) == 43          # 44M1F: 14 codons, then the 4-nt quasi-codon TCaA at target 43..46
    seq_starts = [int(r[9]) for r in rows]
    assert seq_starts == sorted(seq_starts) and 1 <= seq_starts[0] and seq_starts[-1] <= 402
    assert all(int(r[9]) == int(r[10]) for r in rows)          # the hit starts at target position 1: alignment and target positions agree
    del keep


def test_report_without_hits(oracle):
    from bath_b200 import hostapi
    be, keep = oracle.cpu_backend(2)
    model = hostapi.QueryModel(common.golden("PTH2.bhmm"))
    search = hostapi.Search(model, backend=be)
    rng = __import__("numpy").random.default_rng(5)
    search.add_sequence("noise", hostapi.digitize_dna("".join("ACGT"[z] for z in rng.integers(0, 4, 3000))))
    assert search.finish() == []
    text = search.report()
    assert text.count("[No hits detected that satisfy reporting thresholds]") == 2 and text.endswith("\n\n\n")
    del keep


@pytest.mark.gpu
def test_report_matches_shipped_outputs_gpu(gpu_ctx):
    from bath_b200 import hostapi
    check_all(hostapi, gpu_ctx=gpu_ctx)


def test_compare_tables_rules():
    """hostapi.compare_tables (used by bench.py and the large-target parity script): the same hit set by key; a printed float may move by
    one unit of its last digit; hits may change rank only among equal printed E-values; anything else is a difference"""
    from bath_b200 import hostapi
    hdr = "# hit ID  target name ...\n#------- ---\n"
    row = "{:8d} {:<20s} -  q1  PF1  185  {:3d}  183  5950093  {:8d}  {:8d}  {:>8s}  {:5.1f}  {:4.1f} 26.95  {:d}  0 {}\n"
    a = hdr + row.format(1, "contig9", 17, 3125506, 3125006, "1.1e-23", 87.6, 0.1, 0, "501M") \
            + row.format(2, "contig3", 2, 100, 640, "3.2e-14", 60.5, 0.6, 2, "265M2B235M2B") \
            + row.format(3, "contig8", 8, 9000, 8452, "3.2e-14", 60.5, 0.1, 3, "363M1F48M2B4M")
    assert hostapi.compare_tables(a, a) == (True, True, 0)
    b = a.replace(" 87.6   0.1", " 87.6   0.2")                          # a bias digit on a rounding boundary
    assert b != a and hostapi.compare_tables(a, b) == (False, True, 1)
    l = a.splitlines(True)
    swapped = "".join(l[:3] + [l[4].replace("       3 ", "       2 ", 1), l[3].replace("       2 ", "       3 ", 1)])
    assert hostapi.compare_tables(a, swapped) == (False, True, 2)        # equal printed E-values: either order
    assert hostapi.compare_tables(a, a.replace("3125506", "3125507"))[1] is False        # another coordinate: another hit
    assert hostapi.compare_tables(a, a.replace("501M", "500M1I"))[1] is False
    assert hostapi.compare_tables(a, a.replace(" 87.6 ", " 87.9 "))[1] is False          # more than one unit of the last digit
    assert hostapi.compare_tables(a, a.replace("1.1e-23", "1.4e-23"))[1] is False
    assert hostapi.compare_tables(a, "".join(l[:4]))[1] is False                          # a hit missing
    worse = "".join(l[:2] + [l[3].replace("       2 ", "       1 ", 1), l[2].replace("       1 ", "       2 ", 1), l[4]])
    assert hostapi.compare_tables(a, worse)[1] is False                                   # ranks changed across different E-values
