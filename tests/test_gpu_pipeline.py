"""End-to-end parity: the stage-batched pipeline (host C++ + GPU stages) against the reference's golden output
tutorial/AMP_N-fs.out|.tbl: one hit, E-value 1.9e-27, 82.8 bits, bias 0.1, hmm 1-131, target 1-402, 6 frameshifts,
1 stop codon, the CIGAR string, and the filter counters of the footer."""
import re

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def run_search(gpu_ctx, hmm, fasta, **opt):
    from bath_b200 import hostapi
    model = hostapi.QueryModel(common.golden(hmm))
    search = hostapi.Search(model, gpu_ctx, **opt)
    for name, seq in hostapi.read_fasta(common.golden(fasta)):
        search.add_sequence(name, hostapi.digitize_dna(seq))
    hits = search.finish()
    return hits, search.stats()


def test_amp_n_fs_matches_golden_table(gpu_ctx):
    hits, st = run_search(gpu_ctx, "AMP_N.bhmm", "target-AMP_N.fa")
    print(hits, st)
    tbl = [l for l in open(common.golden("AMP_N-fs.tbl")) if not l.startswith("#")]
    assert len(hits) == len(tbl) == 1
    f = tbl[0].split()
    h = hits[0]
    assert h["name"] == f[1]
    assert (h["hmm_from"], h["hmm_to"]) == (int(f[6]), int(f[7]))
    assert h["sq_len"] == int(f[8])
    assert (h["ali_from"], h["ali_to"]) == (int(f[9]), int(f[10]))
    assert f"{h['evalue']:.2g}" == f"{float(f[11]):.2g}"          # 1.9e-27
    assert f"{h['score']:.1f}" == f[12]                            # 82.8
    assert f"{h['bias']:.1f}" == f[13]                             # 0.1
    assert f"{h['pid']:.2f}" == f[14]
    assert (h["shifts"], h["stops"]) == (int(f[15]), int(f[16]))
    assert h["cigar"] == f[17]
    # footer counters of AMP_N-fs.out
    out = open(common.golden("AMP_N-fs.out")).read()
    want = {k: int(re.search(k + r":\s+(\d+)", out).group(1)) for k in
            ("Residues passing SSV filter", "Residues passing bias filter", "Residues passing Vit filter", "Residues passing Fwd filter")}
    assert st["nres"] == 822
    assert (st["pos_past_msv"], st["pos_past_bias"], st["pos_past_vit"], st["pos_past_fwd"]) == tuple(want.values())


def test_pth2_filter_cascade_counters_match_golden(gpu_ctx):
    """tutorial/PTH2.out is a run WITHOUT --fs; the ORF finder and the MSV -> bias -> Viterbi cascade are the same code
    path in both modes (src/p7_pipeline.c:1632-1718), so its footer pins them: 6000 residues searched, 1503 / 1503 / 1401
    residues past the SSV, bias and Viterbi filters."""
    hits, st = run_search(gpu_ctx, "PTH2.bhmm", "target-PTH2.fa")
    out = open(common.golden("PTH2.out")).read()
    want = [int(re.search(k + r":\s+(\d+)", out).group(1)) for k in
            ("Residues passing SSV filter", "Residues passing bias filter", "Residues passing Vit filter")]
    assert st["nres"] == 6000
    assert [st["pos_past_msv"], st["pos_past_bias"], st["pos_past_vit"]] == want
    print(hits, st)


def test_tblout_from_the_gpu_search_is_byte_identical_to_the_shipped_tables(gpu_ctx):
    """header + hit lines of tutorial/AMP_N-fs.tbl (--fs) and tutorial/PTH2.tbl (default pipeline) from the device stages + host pipeline"""
    from bath_b200 import hostapi
    from test_host_pipeline_cpu import golden_table
    for hmm, fasta, tbl, opt in (("AMP_N.bhmm", "target-AMP_N.fa", "AMP_N-fs.tbl", {}),
                                 ("PTH2.bhmm", "target-PTH2.fa", "PTH2.tbl", {"std_only": 1})):
        search = hostapi.Search(hostapi.QueryModel(common.golden(hmm)), gpu_ctx, **opt)
        for name, seq in hostapi.read_fasta(common.golden(fasta)):
            search.add_sequence(name, hostapi.digitize_dna(seq))
        search.finish()
        got = search.tblout()
        search.close()
        assert got == golden_table(tbl), (tbl, got)
