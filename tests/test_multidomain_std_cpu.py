"""The multi-domain branch of the STANDARD-translation domain definition (region_trace_ensemble,
src/p7_domaindef.c:642, :766-860), without a GPU:
 * bath_b200/host/stotrace.cpp's protein flavour (200 tracebacks through the region's Forward matrix reduced to domain end points,
   the per-residue null2 scores taken from the traces, single-linkage clustering) against oracle/stotrace.c (p7_StochasticTrace +
   p7_trace_Index + p7_Null2_ByTrace + p7_spensemble_Cluster on real traces) on the same matrix: identical segments and envelopes,
   null2 scores to 1e-5;
 * the host pipeline behind the CPU backend, default (no --fs) pipeline, on an ORF holding two copies of the homolog back to back:
   the region is flagged multi-domain, resolved into two envelopes, and both copies are reported.
Parity with the reference is UNPINNED for this branch, as for the frameshift flavour (test_multidomain_cpu.py)."""
import ctypes as C

import numpy as np

import common


def partial_copies_target(model, rng, pieces, flank=300):
    """residues sampled from the match states of the node ranges in `pieces`, joined without a gap and back-translated with
    synonymous codons: one ORF whose homologous parts abut with no sharp end between them (a copy that starts in the middle of the
    model right behind a complete one), which is what the region finder cannot cut and is_multidomain_region flags"""
    mat = common.hmm_mat(model)
    codons_for = {a: [] for a in common.AA}
    for idx, a in enumerate(common.STD_CODE):
        if a != "*":
            codons_for[a].append(idx)
    nts, lens = [], []
    for k0, k1 in pieces:
        n0 = len(nts)
        for k in range(k0, k1 + 1):
            p = mat[k].astype(np.float64)
            p /= p.sum()
            c = codons_for[common.AA[rng.choice(20, p=p)]]
            c = c[rng.integers(len(c))]
            nts += [c // 16, (c // 4) % 4, c % 4]
        lens.append(len(nts) - n0)
    return common.embed(rng, np.array(nts, np.uint8), flank, flank), lens


def translate(nt):
    aa = common.AA
    return np.array([aa.index(common.STD_CODE[int(nt[z]) * 16 + int(nt[z + 1]) * 4 + int(nt[z + 2])]) for z in range(0, len(nt) - 2, 3)], np.uint8)


def test_host_sampling_null2_and_clustering_match_oracle(oracle):
    from bath_b200 import hostapi
    po, L = oracle, oracle.lib()
    model = po.Model(common.golden("AMP_N.bhmm"))
    rng = np.random.default_rng(21)
    mat = common.hmm_mat(model)
    for trial, spacer in enumerate((0, 3, 12)):
        a, b = common.sample_homolog(rng, mat, fs_rate=0.0, stop_rate=0.0), common.sample_homolog(rng, mat, fs_rate=0.0, stop_rate=0.0)
        res = np.concatenate([rng.integers(0, 20, 25), translate(a), rng.integers(0, 20, spacer), translate(b), rng.integers(0, 20, 30)]).astype(np.uint8)
        n = len(res)
        dsq = np.concatenate([[255], res, [255]]).astype(np.uint8)
        i, j = 11, n - 9                                            # a region inside the ORF
        Lr = j - i + 1
        om = model.om
        L.bo_oprofile_ReconfigMultihit(om, n)                       # multihit at the ORF's own length (src/p7_domaindef.c:561)
        fwd = L.bo_mx_create(model.M, Lr, 3)
        sc = C.c_float(0)
        sub = np.ascontiguousarray(dsq[i - 1: j + 2]).copy(); sub[0] = 255; sub[-1] = 255
        assert L.bo_Forward(po.u8ptr(sub), Lr, om, fwd, C.byref(sc)) == 0
        cap = 200 * 64
        sp, out = (po.SEGMENT * cap)(), (po.SEGMENT * 64)()
        nsp = C.c_int(0)
        n2 = np.zeros(n + 2, np.float32)
        L.bo_region_trace_ensemble.restype = C.c_int
        L.bo_region_trace_ensemble.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                               C.POINTER(po.SEGMENT), C.c_int, C.POINTER(C.c_int), C.POINTER(po.SEGMENT), C.c_int, C.POINTER(C.c_float)]
        nc = L.bo_region_trace_ensemble(om, po.u8ptr(dsq), fwd, i, j, 42, 200, sp, cap, C.byref(nsp), out, 64, po.fptr(n2))
        assert nc >= 0
        as_t = lambda g: (g.idx, g.i, g.j, g.k, g.m, g.prob)
        want_sp, want_env = [as_t(sp[z]) for z in range(nsp.value)], [as_t(out[z]) for z in range(nc)]

        o = om.contents
        xf = np.ctypeslib.as_array(o.xf).reshape(4, 2)
        odds = [xf[1][0], xf[1][1], xf[0][0], xf[0][1]]             # N move, N loop, E move, E loop
        dp = po.mx_dp(fwd)                                          # [(L+1)][(M+1)][3] M, D, I
        mx4 = np.zeros(dp.shape[:2] + (4,), np.float32); mx4[:, :, :3] = dp
        rbv, rwv, twv, rfv, tfv = model.om_tables()
        got_sp, got_n2 = hostapi.sample_region_segments_protein(mx4, po.mx_xmx(fwd), tfv, rfv, odds, i, res[i - 1: j])
        assert len(want_sp) >= 200 and got_sp == want_sp
        np.testing.assert_allclose(got_n2[1:], n2[i: j + 1], rtol=0, atol=1e-5)
        got_env = hostapi.cluster_region_segments(got_sp, protein=True)
        assert got_env == want_env
        assert len(got_env) == 2, (trial, got_env)
        (_, i1, j1, k1, m1, p1), (_, i2, j2, k2, m2, p2) = got_env
        assert i <= i1 < i2 and j1 < j2 <= j and p1 > 0.5 and p2 > 0.5
        assert k1 < 15 and k2 < 15 and m1 > model.M - 15 and m2 > model.M - 15
        L.bo_mx_destroy(fwd)


def run_std_search(pieces, backend=None, gpu_ctx=None, seed=50):
    from bath_b200 import hostapi
    from oracle import pyoracle as po
    omodel = po.Model(common.golden("AMP_N.bhmm"))
    rng = np.random.default_rng(seed)
    dsq, lens = partial_copies_target(omodel, rng, pieces)
    model = hostapi.QueryModel(common.golden("AMP_N.bhmm"))
    search = hostapi.Search(model, gpu_ctx=gpu_ctx, backend=backend, top_only=1, std_only=1)
    search.add_sequence("tandem", dsq)
    hits = search.finish()
    st, tbl = search.stats(), search.tblout(header=False)
    search.close()
    return hits, st, tbl, lens, omodel.M


STD_CASES = [([(1, 134), (50, 134)], 50, 2), ([(30, 100), (30, 100), (30, 100)], 52, 3)]


def test_default_pipeline_splits_a_multidomain_region(oracle):
    for pieces, seed, ncopies in STD_CASES:
        be, keep = oracle.cpu_backend(4)
        hits, st, tbl, lens, M = run_std_search(pieces, backend=be, seed=seed)
        del keep
        assert st["n_multidomain_regions"] == 1 and st["n_regions"] == 1 and st["n_envelopes"] == ncopies, st
        spans = sorted((h["ali_from"], h["ali_to"], h["hmm_from"], h["hmm_to"]) for h in hits)
        assert len(spans) == ncopies, spans
        at = 301
        for (a, b, k, m), n, (k0, k1) in zip(spans, lens, pieces):
            assert abs(a - at) < 60 and abs(b - (at + n - 1)) < 60, (spans, lens)
            assert abs(k - k0) < 25 and abs(m - k1) < 25, (spans, pieces)
            at += n
        for h in hits:
            assert h["evalue"] < 1e-5
