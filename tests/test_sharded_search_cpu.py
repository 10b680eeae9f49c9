"""The batch / multi-context form of the host pipeline (bathhost_search_create_multi, _queue, _run) on the CPU stage calls:
the merged hit list must not depend on how the target is cut into chunks, how many device contexts the chunks are dealt to, or
whether sequences are searched one at a time or queued and run together -- the reference's serial loop (src/bathsearch.c:1053-1113)
is the definition, whatever the number of workers."""
import numpy as np
import pytest

import common


def _targets(fasta):
    from bath_b200 import hostapi
    return [(name, hostapi.digitize_dna(seq)) for name, seq in hostapi.read_fasta(common.golden(fasta))]


def _search(oracle, hmm, targets, index=0, nbackends=1, queued=True, **opt):
    from bath_b200 import hostapi
    pairs = [oracle.cpu_backend(2) for _ in range(nbackends)]
    model = hostapi.QueryModel(common.golden(hmm), index)
    search = hostapi.Search(model, backend=[p[0] for p in pairs], **opt)
    for name, dsq in targets:
        if queued:
            search.queue_sequence(name, dsq)
        else:
            search.add_sequence(name, dsq)
    hits = search.finish()
    out = (search.tblout(), search.stats(), hits)
    search.close()
    del pairs
    return out


COUNTERS = ("nseqs", "nres", "pos_past_msv", "pos_past_bias", "pos_past_vit", "pos_past_fwd", "n_orfs", "n_windows", "n_std_windows",
            "n_regions", "n_multidomain_regions", "n_envelopes", "n_hits_reported")


def _same(a, b):
    assert a[0] == b[0]
    assert {k: a[1][k] for k in COUNTERS} == {k: b[1][k] for k in COUNTERS}


@pytest.mark.parametrize("fasta", ["2OG-FeII_Oxy_3-nt-fs.fa", "2OG-FeII_Oxy_3-nt.fa"])
@pytest.mark.parametrize("opt", [{}, {"std_only": 1}])
def test_2og_multi_fasta_batch_equals_one_by_one(oracle, fasta, opt):
    """BASELINE config 2: ten target sequences.  One at a time == all queued == dealt to two contexts in tiny chunks."""
    targets = _targets(fasta)
    assert len(targets) == 10
    one = _search(oracle, "2OG-FeII_Oxy_3.bhmm", targets, queued=False, **opt)
    assert one[1]["nseqs"] == 10 and one[1]["n_hits_reported"] >= 8, one[1]
    _same(one, _search(oracle, "2OG-FeII_Oxy_3.bhmm", targets, queued=True, **opt))
    _same(one, _search(oracle, "2OG-FeII_Oxy_3.bhmm", targets, nbackends=2, chunk_nt=1500, **opt))
    _same(one, _search(oracle, "2OG-FeII_Oxy_3.bhmm", targets, nbackends=3, chunk_nt=700, block_length=0, **opt))


def _genome(model, rng, n, ncontigs, every, tandem_at=()):
    from bath_b200 import synth
    out = []
    for c in range(ncontigs):
        dsq, plants = synth.planted_genome(rng, n, model.mat(), every=every, fs_rate=model.fsprob)
        out.append((f"contig{c + 1}", dsq))
    return out


@pytest.mark.parametrize("index", [0, 1, 2])
def test_planted_contigs_are_chunking_and_context_invariant(oracle, index):
    """three tRNA-synthetases models against 3 contigs of 0.4 Mbp with a homolog every 20 kb, small blocks (many block borders,
    overlap duplicates, early E-value cuts on a growing residue count): 1 context / 1 chunk == 2 contexts / 100-kb chunks ==
    sequence by sequence"""
    from bath_b200 import hostapi
    model = hostapi.QueryModel(common.golden("tRNA-synthetases.bhmm"), index)
    rng = np.random.default_rng(100 + index)
    targets = _genome(model, rng, 400_000, 3, 20_000)
    base = _search(oracle, "tRNA-synthetases.bhmm", targets, index=index, block_length=60_000, chunk_nt=10_000_000)
    assert base[1]["n_hits_reported"] >= 40, base[1]
    _same(base, _search(oracle, "tRNA-synthetases.bhmm", targets, index=index, block_length=60_000, nbackends=2, chunk_nt=100_000))
    _same(base, _search(oracle, "tRNA-synthetases.bhmm", targets, index=index, block_length=60_000, queued=False))


def test_finish_twice_changes_nothing_and_more_sequences_can_follow(oracle):
    from bath_b200 import hostapi
    targets = _targets("2OG-FeII_Oxy_3-nt-fs.fa")
    pair = oracle.cpu_backend(2)
    model = hostapi.QueryModel(common.golden("2OG-FeII_Oxy_3.bhmm"))
    search = hostapi.Search(model, backend=pair[0])
    for name, dsq in targets[:5]:
        search.queue_sequence(name, dsq)
    search.finish()
    first = search.tblout()
    search.finish()
    assert search.tblout() == first
    for name, dsq in targets[5:]:
        search.queue_sequence(name, dsq)
    search.finish()
    allten = search.tblout()
    search.close()
    assert allten == _search(oracle, "2OG-FeII_Oxy_3.bhmm", targets)[0]


def test_three_profiles_finished_at_once_equal_one_after_the_other(oracle):
    """bathhost_search_finish_many: the query file's three profiles against the same contigs at the same time (one host thread per
    search, two contexts each) give, search by search, the tables of the reference's one-query-at-a-time loop (src/bathsearch.c:737)"""
    from bath_b200 import hostapi
    models = [hostapi.QueryModel(common.golden("tRNA-synthetases.bhmm"), i) for i in range(3)]
    rng = np.random.default_rng(7)
    targets = []
    for c in range(3):
        from bath_b200 import synth
        dsq, _ = synth.planted_genome(rng, 200_000, models[c].mat(), every=20_000, fs_rate=models[c].fsprob)
        targets.append((f"contig{c + 1}", dsq))
    alone = [_search(oracle, "tRNA-synthetases.bhmm", targets, index=i, block_length=60_000) for i in range(3)]
    pairs = [[oracle.cpu_backend(2) for _ in range(2)] for _ in models]
    searches = [hostapi.Search(m, backend=[p[0] for p in ps], block_length=60_000, chunk_nt=150_000) for m, ps in zip(models, pairs)]
    for s in searches:
        for name, dsq in targets:
            s.queue_sequence(name, dsq)
    hostapi.Search.finish_many(searches)
    for s, ref in zip(searches, alone):
        assert ref[1]["n_hits_reported"] >= 10
        _same((s.tblout(), s.stats(), None), ref)
    # a search twice, or two searches over one context, are refused
    lib = searches[0].lib
    import ctypes as C
    arr = (C.c_void_p * 2)(searches[0].h, searches[0].h)
    assert lib.bathhost_search_finish_many(arr, 2) == 11          # BATHHOST_EINVAL
    shared = hostapi.Search(models[1], backend=[pairs[0][0][0]])
    arr = (C.c_void_p * 2)(searches[0].h, shared.h)
    assert lib.bathhost_search_finish_many(arr, 2) == 11
    assert lib.bathhost_search_finish_many(None, 0) == hostapi.OK
    shared.close()
    for s in searches:
        s.close()
    del pairs


def test_segment_upload_equals_host_concatenation(oracle):
    """a chunk made of several sequences: handed to the backend piece by piece (upload_block_segments) or concatenated on the host
    first (a backend without the entry) -- the same tables"""
    from bath_b200 import hostapi
    targets = _targets("2OG-FeII_Oxy_3-nt-fs.fa")
    model = hostapi.QueryModel(common.golden("2OG-FeII_Oxy_3.bhmm"))
    out = []
    for drop in (False, True):
        pair = oracle.cpu_backend(2)
        if drop:
            pair[0].upload_block_segments = None
        search = hostapi.Search(model, backend=pair[0], chunk_nt=10_000_000)      # all ten sequences in one chunk
        for name, dsq in targets:
            search.queue_sequence(name, dsq)
        search.finish()
        out.append((search.tblout(), search.stats()))
        search.close()
        del pair
    _same(out[0], out[1])
    assert out[0][1]["n_hits_reported"] >= 8
