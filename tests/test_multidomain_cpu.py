"""The multi-domain branch of frameshift domain definition, without a GPU:
 * bath_b200/host/stotrace.cpp (sampling of 200 tracebacks reduced to domain end points, single-linkage clustering, dominated
   clusters) against oracle/fs_stotrace.c (p7_StochasticTrace_Frameshift + p7_trace_fs_Index + p7_spensemble_fs_Cluster) on the same
   Forward matrix: identical segments, identical envelopes;
 * the host pipeline behind the CPU backend on a target with two homologs back to back: the region is flagged multi-domain, split
   by the clustering, and both copies are reported as hits of their own.
Parity with the reference is UNPINNED for this branch (no shipped output exercises it; the generator and the clustering routine are
Easel's, restated): these tests pin the product against the oracle restatement only."""
import ctypes as C

import numpy as np

import common


def tandem_target(oracle, model, rng, spacer=6, flank=400, copies=2):
    mat = common.hmm_mat(model)
    parts = []
    for c in range(copies):
        parts.append(common.sample_homolog(rng, mat, fs_rate=0.01, stop_rate=0.0))
        if c + 1 < copies:
            parts.append(rng.integers(0, 4, spacer).astype(np.uint8))
    ins = np.concatenate(parts)
    return common.embed(rng, ins, flank, flank), [len(p) for p in parts]


def multihit_forward(oracle, model, dsq, i, j):
    """the reference's call at src/p7_domaindef.c:409-412: om_fs5 reconfigured multihit for length 100, Forward over dsq[i..j]"""
    L = oracle.lib()
    om = model.om_fs5
    L.bo_fs_oprofile_ReconfigMultihit(om, 100)
    Lr = j - i + 1
    sub = np.full(Lr + 2, 255, np.uint8)
    sub[1:-1] = dsq[i:j + 1]
    fwd = L.bo_mx_create(model.M, Lr, 8)
    sc = C.c_float(0)
    st = L.bo_Forward_Frameshift(oracle.u8ptr(sub), Lr, om, fwd, C.byref(sc))
    assert st == 0
    return fwd, sc.value


def test_host_sampling_and_clustering_match_oracle(oracle):
    from bath_b200 import hostapi
    model = oracle.Model(common.golden("AMP_N.bhmm"))
    rng = np.random.default_rng(5)
    for trial, spacer in enumerate((0, 6, 45)):
        dsq, parts = tandem_target(oracle, model, rng, spacer=spacer, flank=60)
        n = len(dsq) - 2
        i, j = 31, n - 20                                   # a region inside the target, as the region finder would cut it
        fwd, _ = multihit_forward(oracle, model, dsq, i, j)
        want_sp, want_env = oracle.region_trace_ensemble(model.om_fs5, fwd, i, j)
        xf = model.xf(5)                                    # [E,N,J,C][MOVE,LOOP]
        odds = [xf[1][0], xf[1][1], xf[0][0], xf[0][1]]
        assert abs(odds[0] - 3.0 / 103.0) < 1e-7 and odds[2] == 0.5 and odds[3] == 0.5
        got_sp = hostapi.sample_region_segments(oracle.mx_dp(fwd), oracle.mx_xmx(fwd), model.tfv(5), odds, i)
        assert len(want_sp) >= 200 and got_sp == want_sp
        got_env = hostapi.cluster_region_segments(got_sp)
        assert got_env == want_env
        # two homologs back to back: two envelopes, in order, each about one model long and inside the region
        assert len(got_env) == 2, (trial, got_env)
        (_, i1, j1, k1, m1, p1), (_, i2, j2, k2, m2, p2) = got_env
        assert i <= i1 < i2 and j1 < j2 <= j and j1 - i2 < 30 and p1 > 0.5 and p2 > 0.5      # consensus end points may overlap by a few codons
        assert k1 < 15 and k2 < 15 and m1 > model.M - 15 and m2 > model.M - 15
        oracle.lib().bo_mx_destroy(fwd)
    oracle.lib().bo_fs_oprofile_ReconfigUnihit(model.om_fs5, 100)


def test_clustering_edge_cases():
    from bath_b200 import hostapi
    assert hostapi.cluster_region_segments([]) == []
    # one trace in four carries the domain: posterior 0.25 is kept, below it the cluster is dropped
    seg = [(t, 100, 400, 1, 100, 0.0) for t in range(0, 200, 4)]
    out = hostapi.cluster_region_segments(seg)
    assert len(out) == 1 and out[0][1:5] == (100, 400, 1, 100) and abs(out[0][5] - 0.25) < 1e-7
    assert hostapi.cluster_region_segments(seg[:-1]) == []
    # end points: the widest one carried by >= 2 % of the cluster's traces wins; a single outlier among 200 does not
    seg = [(t, 100 + (t % 3), 400, 5, 100, 0.0) for t in range(200)]
    seg[7] = (7, 40, 400, 5, 100, 0.0)
    out = hostapi.cluster_region_segments(seg)
    assert len(out) == 1 and out[0][1] == 100 and out[0][2] == 400
    # two domains per trace far apart on the sequence: two clusters ordered by start; a weaker cluster overlapping >= 80 % of a
    # stronger one on the sequence (other end of the model) is dominated and removed
    seg = []
    for t in range(200):
        seg.append((t, 100, 400, 1, 100, 0.0))
        seg.append((t, 700, 1000, 1, 100, 0.0))
        if t % 2 == 0:
            seg.append((t, 705, 1000, 60, 160, 0.0))
    out = hostapi.cluster_region_segments(seg)
    assert [g[1:3] for g in out] == [(100, 400), (700, 1000)]


def test_pipeline_splits_a_multidomain_region(oracle):
    """two copies of the AMP_N homolog back to back: one merged window, one region, flagged multi-domain, two envelopes, two hits"""
    from bath_b200 import hostapi
    omodel = oracle.Model(common.golden("AMP_N.bhmm"))
    rng = np.random.default_rng(11)
    dsq, parts = tandem_target(oracle, omodel, rng, spacer=0, flank=500)
    be, keep = oracle.cpu_backend(4)
    model = hostapi.QueryModel(common.golden("AMP_N.bhmm"))
    search = hostapi.Search(model, backend=be, top_only=1)
    search.add_sequence("tandem", dsq)
    hits = search.finish()
    st = search.stats()
    search.close()
    del keep
    assert st["n_multidomain_regions"] == 1 and st["n_regions"] == 1 and st["n_envelopes"] == 2
    spans = sorted((h["ali_from"], h["ali_to"]) for h in hits)
    a0 = 501
    a1 = a0 + parts[0] + parts[1]
    assert len(spans) == 2, spans
    assert abs(spans[0][0] - a0) < 45 and abs(spans[0][1] - (a0 + parts[0] - 1)) < 45
    assert abs(spans[1][0] - a1) < 45 and abs(spans[1][1] - (a1 + parts[2] - 1)) < 45
    for h in hits:
        assert h["hmm_from"] < 15 and h["hmm_to"] > omodel.M - 15 and h["evalue"] < 1e-10
