"""Hit-set identity of the device pipeline on the targets the golden outputs do not reach:
BASELINE config 2 (2OG-FeII_Oxy_3.bhmm against its ten-sequence nt / nt-fs targets), multi-Mbp planted genomes with all three
tRNA-synthetases models and PTHR37536, default and small blocks -- the GPU search (one context, two contexts on one device, one
context per device when the box has two) must write the SAME --tblout table, byte for byte, as the same host pipeline over the CPU
oracle's stage calls, and the same table whatever the number of contexts."""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu

COUNTERS = ("nseqs", "nres", "pos_past_msv", "pos_past_bias", "pos_past_vit", "pos_past_fwd", "n_orfs", "n_windows", "n_std_windows",
            "n_regions", "n_multidomain_regions", "n_envelopes", "n_hits_reported")


def _run(model, targets, gpu=None, backend=None, **opt):
    from bath_b200 import hostapi
    search = hostapi.Search(model, gpu_ctx=gpu, backend=backend, **opt)
    for name, dsq in targets:
        search.queue_sequence(name, dsq)
    search.finish()
    out = (search.tblout(), {k: v for k, v in search.stats().items() if k in COUNTERS})
    search.close()
    return out


def _cpu(oracle, model, targets, **opt):
    be, keep = oracle.cpu_backend(16)
    out = _run(model, targets, backend=be, **opt)
    del keep
    return out


@pytest.fixture(scope="module")
def contexts():
    """two contexts on device 0, plus one on device 1 when there is one"""
    import torch
    from bath_b200 import capi
    ctxs = [capi.Context(0), capi.Context(0)]
    if torch.cuda.device_count() > 1:
        ctxs.append(capi.Context(1))
    yield ctxs
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("fasta", ["2OG-FeII_Oxy_3-nt-fs.fa", "2OG-FeII_Oxy_3-nt.fa"])
@pytest.mark.parametrize("opt", [{}, {"std_only": 1}])
def test_config2_2og_gpu_equals_cpu_backend(oracle, contexts, fasta, opt):
    from bath_b200 import hostapi
    model = hostapi.QueryModel(common.golden("2OG-FeII_Oxy_3.bhmm"))
    targets = [(n, hostapi.digitize_dna(s)) for n, s in hostapi.read_fasta(common.golden(fasta))]
    want = _cpu(oracle, model, targets, **opt)
    assert want[1]["nseqs"] == 10 and want[1]["n_hits_reported"] >= 8
    assert _run(model, targets, gpu=contexts[0], **opt) == want
    assert _run(model, targets, gpu=contexts, chunk_nt=1500, **opt) == want


def _contigs(model, seed, sizes, every, tandem=True):
    from bath_b200 import synth
    rng = np.random.default_rng(seed)
    out = []
    for c, n in enumerate(sizes):
        dsq, plants = synth.planted_genome(rng, n, model.mat(), every=every, fs_rate=model.fsprob)
        if tandem and len(plants) > 3:                         # two homologs back to back: a multi-domain region
            a, b, strand = plants[1]
            h = dsq[a:b + 1].copy()
            if b + 1 + len(h) + 30 < n:
                dsq[b + 31:b + 31 + len(h)] = h
        out.append((f"contig{c + 1}", dsq))
    return out


@pytest.mark.parametrize("hmm,index", [("tRNA-synthetases.bhmm", 0), ("tRNA-synthetases.bhmm", 1), ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0)])
def test_planted_genome_12mbp_gpu_equals_cpu_backend(oracle, contexts, hmm, index):
    """12 Mbp in four contigs, a homolog every 20 kb (some tandem), default block length: one context == several == CPU backend"""
    from bath_b200 import hostapi
    model = hostapi.QueryModel(common.golden(hmm), index)
    targets = _contigs(model, 7 + index, [5_000_000, 3_000_000, 2_500_000, 1_500_000], 20_000)
    want = _cpu(oracle, model, targets)
    assert want[1]["n_hits_reported"] >= 500, want[1]
    one = _run(model, targets, gpu=contexts[0])
    assert one[1] == want[1]
    assert one[0] == want[0]
    assert _run(model, targets, gpu=contexts, chunk_nt=1_000_000) == want


def test_small_blocks_gpu_equals_cpu_backend(oracle, contexts):
    """40-kb blocks over 3 Mbp: hundreds of block borders with overlap duplicates and the growing residue count of the early cuts"""
    from bath_b200 import hostapi
    model = hostapi.QueryModel(common.golden("tRNA-synthetases.bhmm"), 1)
    targets = _contigs(model, 99, [2_000_000, 1_000_000], 15_000)
    want = _cpu(oracle, model, targets, block_length=40_000)
    assert _run(model, targets, gpu=contexts[0], block_length=40_000) == want
    assert _run(model, targets, gpu=contexts, block_length=40_000, chunk_nt=300_000) == want


def test_three_profiles_at_once_equal_one_after_the_other(oracle):
    """bathhost_search_finish_many on the device: the three tRNA-synthetases profiles against the same 6 Mbp at the same time, two
    contexts each, give the tables of the one-profile-at-a-time loop and of the CPU backend"""
    from bath_b200 import capi, hostapi, synth
    models = [hostapi.QueryModel(common.golden("tRNA-synthetases.bhmm"), i) for i in range(3)]
    rng = np.random.default_rng(11)
    contigs, _ = synth.planted_contigs(rng, 6_000_000, [m.mat() for m in models], every=25_000, fs_rates=[m.fsprob for m in models],
                                       min_len=1_000_000, max_len=3_000_000)
    sets = [[capi.Context(0), capi.Context(0)] for _ in models]
    alone = [_run(m, contigs, gpu=cs) for m, cs in zip(models, sets)]
    searches = [hostapi.Search(m, gpu_ctx=cs, chunk_nt=1_500_000) for m, cs in zip(models, sets)]
    for s in searches:
        for name, dsq in contigs:
            s.queue_sequence(name, dsq)
    hostapi.Search.finish_many(searches)
    for s, ref, m in zip(searches, alone, models):
        got = (s.tblout(), {k: v for k, v in s.stats().items() if k in COUNTERS})
        assert ref[1]["n_hits_reported"] >= 60, ref[1]
        assert got == ref
        cpu = _cpu(oracle, m, contigs)
        # against the CPU backend: the same hits, coordinates and CIGAR strings; a printed float may sit on a rounding boundary
        # (here one E-value of the third profile prints 6.2e-31 / 6.3e-31: ln P differs by < 5e-4)
        same_bytes, equivalent, ndiff = hostapi.compare_tables(got[0], cpu[0])
        assert equivalent and ndiff <= 1 and got[1] == cpu[1]
        s.close()
    for cs in sets:
        for c in cs:
            c.close()
