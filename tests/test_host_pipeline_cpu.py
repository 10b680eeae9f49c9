"""The host pipeline (bath_b200/host/pipeline.cpp) behind the CPU oracle's implementation of the stage calls
(oracle/cpu_backend.c): end-to-end parity with the reference's golden outputs WITHOUT a GPU.  The same assertions run
against the GPU stages in tests/test_gpu_pipeline.py."""
import re

import common


def run_search_cpu(oracle, hmm, fasta, nthreads=4, **opt):
    from bath_b200 import hostapi
    be, keep = oracle.cpu_backend(nthreads)
    model = hostapi.QueryModel(common.golden(hmm))
    search = hostapi.Search(model, backend=be, **opt)
    for name, seq in hostapi.read_fasta(common.golden(fasta)):
        search.add_sequence(name, hostapi.digitize_dna(seq))
    hits = search.finish()
    st = search.stats()
    search.close()
    del keep
    return hits, st


def footer_counts(path, keys):
    out = open(path).read()
    return [int(re.search(k + r":\s+(\d+)", out).group(1)) for k in keys]


def test_amp_n_fs_table_and_footer(oracle):
    hits, st = run_search_cpu(oracle, "AMP_N.bhmm", "target-AMP_N.fa")
    tbl = [l for l in open(common.golden("AMP_N-fs.tbl")) if not l.startswith("#")]
    assert len(hits) == len(tbl) == 1
    f, h = tbl[0].split(), hits[0]
    assert h["name"] == f[1] and h["sq_len"] == int(f[8])
    assert (h["hmm_from"], h["hmm_to"], h["ali_from"], h["ali_to"]) == (int(f[6]), int(f[7]), int(f[9]), int(f[10]))
    assert f"{h['evalue']:.2g}" == f"{float(f[11]):.2g}"          # 1.9e-27
    assert f"{h['score']:.1f}" == f[12] and f"{h['bias']:.1f}" == f[13] and f"{h['pid']:.2f}" == f[14]
    assert (h["shifts"], h["stops"]) == (int(f[15]), int(f[16])) and h["cigar"] == f[17]
    want = footer_counts(common.golden("AMP_N-fs.out"), ("Residues passing SSV filter", "Residues passing bias filter",
                                                          "Residues passing Vit filter", "Residues passing Fwd filter"))
    assert st["nres"] == 822
    assert [st["pos_past_msv"], st["pos_past_bias"], st["pos_past_vit"], st["pos_past_fwd"]] == want


def test_pth2_filter_counters(oracle):
    hits, st = run_search_cpu(oracle, "PTH2.bhmm", "target-PTH2.fa")
    want = footer_counts(common.golden("PTH2.out"), ("Residues passing SSV filter", "Residues passing bias filter",
                                                      "Residues passing Vit filter"))
    assert st["nres"] == 6000
    assert [st["pos_past_msv"], st["pos_past_bias"], st["pos_past_vit"]] == want
    # the frameshift-aware run finds the non-frameshift run's alignments 2 and 4 (tutorial/PTH2.tbl) as well
    spans = {(min(h["ali_from"], h["ali_to"]), max(h["ali_from"], h["ali_to"])) for h in hits}
    assert any(abs(a - 1486) <= 3 and abs(b - 1731) <= 3 for a, b in spans)
    assert any(abs(a - 1273) <= 3 for a, b in spans)


def test_strand_restriction_and_blocks(oracle):
    """top-only / bottom-only searches partition the hits; a tiny block length (many blocks with overlap context) finds the same
    alignments after duplicate removal"""
    both, st = run_search_cpu(oracle, "PTH2.bhmm", "target-PTH2.fa")
    top, _ = run_search_cpu(oracle, "PTH2.bhmm", "target-PTH2.fa", top_only=1)
    bot, _ = run_search_cpu(oracle, "PTH2.bhmm", "target-PTH2.fa", bottom_only=1)
    assert all(h["strand"] == 1 for h in top) and all(h["strand"] == -1 for h in bot)
    assert len(top) + len(bot) == len(both)
    small, st2 = run_search_cpu(oracle, "PTH2.bhmm", "target-PTH2.fa", block_length=1500)
    assert st2["nres"] == st["nres"]
    # a block boundary may cut a window short (the reference keeps the better of the overlapping duplicates,
    # src/p7_tophits.c:816-900), so alignments are matched by overlap on the target, not by exact ends
    def span(h):
        return h["strand"], min(h["ali_from"], h["ali_to"]), max(h["ali_from"], h["ali_to"])
    assert len(both) <= len(small) <= len(both) + 1       # a window cut by a block boundary can surface a second, partly overlapping alignment
    for h in both:
        s0, a0, b0 = span(h)
        assert any(s1 == s0 and a1 <= b0 and b1 >= a0 for s1, a1, b1 in map(span, small)), h


def _row(h):
    return (h["hmm_from"], h["hmm_to"], h["ali_from"], h["ali_to"], f"{h['evalue']:.2g}", f"{h['score']:.1f}", f"{h['bias']:.1f}")


def check_pth2_default(hits, st):
    """tutorial/PTH2.tbl + PTH2.out: bathsearch WITHOUT --fs, i.e. the standard-translation branch alone (ORF Forward/Backward,
    domain decoding, envelope rescoring, optimal-accuracy alignment mapped back to nucleotides)"""
    tbl = [l.split() for l in open(common.golden("PTH2.tbl")) if not l.startswith("#")]
    assert len(hits) == len(tbl) == 4
    for f, h in zip(tbl, hits):
        assert _row(h) == (int(f[6]), int(f[7]), int(f[9]), int(f[10]), f"{float(f[11]):.2g}", f[12], f[13]), (f, h)
        assert f"{h['pid']:.2f}" == f[14] and h["cigar"] == f[15] and h["shifts"] == 0
    want = footer_counts(common.golden("PTH2.out"), ("Residues passing SSV filter", "Residues passing bias filter",
                                                      "Residues passing Vit filter", "Residues passing Fwd filter"))
    assert [st["pos_past_msv"], st["pos_past_bias"], st["pos_past_vit"], st["pos_past_fwd"]] == want


def check_amp_n_default(hits, st):
    """tutorial/AMP_N.out (no --fs): one hit, 47.8 bits, E 1.4e-16, hmm 3-75, ali 7-234, env 1-237; 237 residues past Forward"""
    out = open(common.golden("AMP_N.out")).read()
    f = re.search(r"^ !\s+(\S+)\s+(\S+)\s+(\S+)\s+(\d+)\s+(\d+) \S\S\s+(\d+)\s+(\d+) \S\S\s+(\d+)\s+(\d+)", out, re.M).groups()
    assert len(hits) == 1
    h = hits[0]
    assert (f"{h['score']:.1f}", f"{h['bias']:.1f}", f"{h['evalue']:.2g}") == (f[0], f[1], f"{float(f[2]):.2g}")
    assert (h["hmm_from"], h["hmm_to"], h["ali_from"], h["ali_to"], h["env_from"], h["env_to"]) == tuple(int(x) for x in f[3:9])
    assert st["pos_past_fwd"] == footer_counts(common.golden("AMP_N.out"), ("Residues passing Fwd filter",))[0]


def test_default_pipeline_pth2_and_amp_n(oracle):
    check_pth2_default(*run_search_cpu(oracle, "PTH2.bhmm", "target-PTH2.fa", std_only=1))
    check_amp_n_default(*run_search_cpu(oracle, "AMP_N.bhmm", "target-AMP_N.fa", std_only=1))


def check_met_ct4(make_search):
    """tutorial/MET-ct4.out: two queries (M = 409, 458) built for codon table 4 against a 35.6 kb target, both strands:
    every hit line (score, bias, E-value, model / alignment / envelope coordinates) and the four filter counters"""
    from bath_b200 import hostapi
    out = open(common.golden("MET-ct4.out")).read()
    parts = out.split("Query:       ")[1:]
    for idx, part in enumerate(parts):
        rows = re.findall(r"^ !\s+(\S+)\s+(\S+)\s+(\S+)\s+(\d+)\s+(\d+) \S\S\s+(\d+)\s+(\d+) \S\S\s+(\d+)\s+(\d+)", part, re.M)
        model = hostapi.QueryModel(common.golden("MET-ct4.bhmm"), index=idx)
        search = make_search(model)
        for name, seq in hostapi.read_fasta(common.golden("target-MET.fa")):
            search.add_sequence(name, hostapi.digitize_dna(seq))
        hits, st = search.finish(), search.stats()
        search.close()
        assert len(hits) == len(rows) == 3
        for f, h in zip(rows, hits):
            assert (f"{h['score']:.1f}", f"{h['bias']:.1f}", f"{h['evalue']:.2g}") == (f[0], f[1], f"{float(f[2]):.2g}")
            assert (h["hmm_from"], h["hmm_to"], h["ali_from"], h["ali_to"], h["env_from"], h["env_to"]) == tuple(int(x) for x in f[3:9])
        want = [int(re.search(k + r":\s+(\d+)", part).group(1)) for k in
                ("Residues passing SSV filter", "Residues passing bias filter", "Residues passing Vit filter", "Residues passing Fwd filter")]
        assert [st["pos_past_msv"], st["pos_past_bias"], st["pos_past_vit"], st["pos_past_fwd"]] == want
        assert st["nres"] == 71226


def test_default_pipeline_met_codon_table_4(oracle):
    from bath_b200 import hostapi
    be, keep = oracle.cpu_backend(4)
    check_met_ct4(lambda model: hostapi.Search(model, backend=be, std_only=1))
    del keep


def test_fs_pipeline_keeps_unshifted_hits(oracle):
    """--fs on PTH2: the window whose ORF beats the frameshift Forward score goes down the standard-translation branch and comes
    out exactly as in the default pipeline (tutorial/PTH2.tbl hit 1); the others are found by the frameshift branch"""
    hits, st = run_search_cpu(oracle, "PTH2.bhmm", "target-PTH2.fa")
    f = [l.split() for l in open(common.golden("PTH2.tbl")) if not l.startswith("#")][0]
    assert st["n_std_windows"] == 1 and len(hits) == 4
    assert _row(hits[0]) == (int(f[6]), int(f[7]), int(f[9]), int(f[10]), f"{float(f[11]):.2g}", f[12], f[13]) and hits[0]["cigar"] == f[15]


def test_speculative_region_walk_equals_the_sequential_walk(oracle):
    """The region walk chains the length model from window to window (src/p7_domaindef.c:320-325, :1018).  The host pipeline walks
    all windows in parallel from guessed inputs and repeats what started from a wrong one; BATHHOST_SEQUENTIAL_WALK forces one
    window per round from its true input.  Both must give the same hits, envelope for envelope, on a target with many windows
    (planted homologs, some back to back so that the multi-domain branch is in the chain as well)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, json, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import common
from oracle import pyoracle as oracle
from bath_b200 import hostapi
om = oracle.Model(common.golden("AMP_N.bhmm"))
mat = common.hmm_mat(om)
rng = np.random.default_rng(3)
parts = []
for t in range(14):
    parts.append(rng.integers(0, 4, int(rng.integers(300, 2500))).astype(np.uint8))
    parts.append(common.sample_homolog(rng, mat, fs_rate=0.02, stop_rate=0.002))
    if t %% 4 == 1:
        parts.append(common.sample_homolog(rng, mat, fs_rate=0.01, stop_rate=0.0))
body = np.concatenate(parts)
dsq = np.full(len(body) + 2, 255, np.uint8); dsq[1:-1] = body
be, keep = oracle.cpu_backend(4)
search = hostapi.Search(hostapi.QueryModel(common.golden("AMP_N.bhmm")), backend=be, block_length=4000)
search.add_sequence("many", dsq)
hits = search.finish(); st = search.stats(); search.close()
print(json.dumps({"hits": [[h["ali_from"], h["ali_to"], h["env_from"], h["env_to"], h["hmm_from"], h["hmm_to"], h["cigar"], "%%.6g" %% h["evalue"]] for h in hits],
                  "stats": {k: st[k] for k in ("n_windows", "n_regions", "n_multidomain_regions", "n_envelopes")}}))
''' % (root, os.path.join(root, "tests"))
    runs = []
    for seq in (False, True):
        env = dict(os.environ)
        env.pop("BATHHOST_SEQUENTIAL_WALK", None)
        if seq:
            env["BATHHOST_SEQUENTIAL_WALK"] = "1"
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        runs.append(__import__("json").loads(out.stdout.strip().splitlines()[-1]))
    assert runs[0] == runs[1]
    assert runs[0]["stats"]["n_windows"] >= 10 and runs[0]["stats"]["n_multidomain_regions"] >= 1 and len(runs[0]["hits"]) >= 14


def golden_table(name):
    """header and hit lines of a shipped --tblout file (everything before the trailer that starts with a bare '#')"""
    lines = open(common.golden(name)).read().split("\n")
    return "\n".join(lines[:lines.index("#")]) + "\n"


def test_tblout_is_byte_identical_to_the_shipped_tables(oracle):
    """bathhost_search_format_tblout (p7_tophits_TabularTargets, src/p7_tophits.c:1603-1712) against tutorial/AMP_N-fs.tbl
    (bathsearch --fs --cigar) and tutorial/PTH2.tbl (default pipeline, --cigar): header and every hit line, byte for byte."""
    from bath_b200 import hostapi
    for hmm, fasta, tbl, opt in (("AMP_N.bhmm", "target-AMP_N.fa", "AMP_N-fs.tbl", {}),
                                 ("PTH2.bhmm", "target-PTH2.fa", "PTH2.tbl", {"std_only": 1})):
        be, keep = oracle.cpu_backend(4)
        search = hostapi.Search(hostapi.QueryModel(common.golden(hmm)), backend=be, **opt)
        for name, seq in hostapi.read_fasta(common.golden(fasta)):
            search.add_sequence(name, hostapi.digitize_dna(seq))
        search.finish()
        got = search.tblout()
        search.close()
        del keep
        assert got == golden_table(tbl), (tbl, got)
