"""GPU parity for the multi-domain branch through the C ABI:
 * bathgpu_fs_forward_matrices (p7_Forward_Frameshift with the full matrix kept, D cells included, multihit configuration) against
   the oracle's matrix: every cell and every X-row entry within 2e-4 of the row's largest cell (rows are rescaled to O(1); observed
   1e-6), scores within 1e-3 nat, over several models incl. the 16/24-nodes-per-lane kernels;
 * the host's sampling and clustering run on the DEVICE's matrix give the oracle's envelopes;
 * the whole search (GPU stages + host pipeline) on two homologs back to back reports the same two hits as the CPU backend.
Parity with the reference for this branch is unpinned (see tests/test_multidomain_cpu.py)."""
import ctypes as C

import numpy as np
import pytest

import common
from test_multidomain_cpu import multihit_forward, tandem_target

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("tRNA-synthetases.bhmm", 1), ("PTHR37536.bhmm", 0), ("MET-ct4.bhmm", 0),
                                           ("synthetic_M624.bhmm", 0)])
def test_forward_matrices_match_oracle(oracle, gpu_ctx, hmmfile, index):
    from bath_b200 import capi, hostapi
    model = oracle.Model(common.golden(hmmfile), index)
    rng = np.random.default_rng(3 + index)
    dsq, parts = tandem_target(oracle, model, rng, spacer=0, flank=150)
    n = len(dsq) - 2
    regions = [(101, n - 90), (1, n), (140, 140 + parts[0] // 2)]          # the tandem pair, the whole block, half a homolog
    gpu_ctx.load_fs_profile(5, model.rfv(5), model.tfv(5))
    gpu_ctx.upload_block(dsq)
    pm, pl = hostapi.length_model(100, nj=1.0)
    regs = capi.Context.make_windows([i for i, _ in regions], [j - i + 1 for i, j in regions], nj=1.0)
    regs["pmove"] = pm; regs["ploop"] = pl
    mx, xr, off, sc, st = gpu_ctx.fs_forward_matrices(regs, model.M, xfE5=(0.5, 0.5))
    worst = 0.0
    for r, (i, j) in enumerate(regions):
        fwd, want_sc = multihit_forward(oracle, model, dsq, i, j)
        assert st[r] == 0 and abs(sc[r] - want_sc) <= 1e-3, (r, sc[r], want_sc)
        L = j - i + 1
        g_mx, g_xr = mx[off[r]: off[r] + L + 1], xr[off[r]: off[r] + L + 1]
        w_mx, w_xr = oracle.mx_dp(fwd), oracle.mx_xmx(fwd)
        assert np.array_equal(g_xr[:, 5], w_xr[:, 5]) or np.allclose(g_xr[:, 5], w_xr[:, 5], rtol=1e-5), "scale factors"
        ref = np.maximum(w_mx.reshape(L + 1, -1).max(axis=1), w_xr[:, :5].max(axis=1))       # largest entry of each row
        ref = np.maximum(ref, 1e-30)[:, None]
        d = max(float(np.max(np.abs(g_mx - w_mx).reshape(L + 1, -1) / ref)), float(np.max(np.abs(g_xr[:, :5] - w_xr[:, :5]) / ref)))
        assert d <= 2e-4, (hmmfile, r, d)
        worst = max(worst, d)
        if r == 0:
            # the host's sampling and clustering on the device's matrix: the oracle's envelopes
            want_sp, want_env = oracle.region_trace_ensemble(model.om_fs5, fwd, i, j)
            xf = model.xf(5)
            got_sp = hostapi.sample_region_segments(g_mx, g_xr, model.tfv(5), [xf[1][0], xf[1][1], xf[0][0], xf[0][1]], i)
            got_env = hostapi.cluster_region_segments(got_sp)
            same = sum(1 for a, b in zip(got_sp, want_sp) if a == b)
            print(f"{hmmfile}: {len(want_sp)} sampled segments, {same} identical on the device matrix; envelopes {got_env}")
            assert len(got_env) == len(want_env) == 2
            for g, w in zip(got_env, want_env):
                assert all(abs(a - b) <= 6 for a, b in zip(g[1:5], w[1:5])), (g, w)
        oracle.lib().bo_mx_destroy(fwd)
    oracle.lib().bo_fs_oprofile_ReconfigUnihit(model.om_fs5, 100)
    print(f"{hmmfile}[{index}] M={model.M}: worst relative cell error {worst:.2e}")


def test_forward_matrices_bad_arguments(gpu_ctx, oracle):
    from bath_b200 import capi
    model = oracle.Model(common.golden("AMP_N.bhmm"))
    gpu_ctx.load_fs_profile(5, model.rfv(5), model.tfv(5))
    gpu_ctx.upload_block(common.random_dna(np.random.default_rng(1), 500))
    regs = capi.Context.make_windows([400], [200], nj=1.0)               # runs past the block
    with pytest.raises(RuntimeError):
        gpu_ctx.fs_forward_matrices(regs, model.M)


def test_search_splits_a_multidomain_region_like_the_cpu_backend(oracle, gpu_ctx):
    from bath_b200 import hostapi
    omodel = oracle.Model(common.golden("AMP_N.bhmm"))
    results = []
    for seed in (11, 12):
        rng = np.random.default_rng(seed)
        dsq, parts = tandem_target(oracle, omodel, rng, spacer=0, flank=500)
        runs = []
        for backend in ("cpu", "gpu"):
            model = hostapi.QueryModel(common.golden("AMP_N.bhmm"))
            if backend == "cpu":
                be, keep = oracle.cpu_backend(4)
                search = hostapi.Search(model, backend=be)
            else:
                search = hostapi.Search(model, gpu_ctx)
            search.add_sequence("tandem", dsq)
            hits = search.finish()
            st = search.stats()
            search.close()
            runs.append((hits, st))
        (ch, cs), (gh, gs) = runs
        assert cs["n_multidomain_regions"] == gs["n_multidomain_regions"] >= 1
        assert cs["n_envelopes"] == gs["n_envelopes"] and len(ch) == len(gh) == 2
        key = lambda h: (h["ali_from"], h["ali_to"], h["hmm_from"], h["hmm_to"], h["cigar"], f"{h['score']:.1f}", f"{h['evalue']:.2g}")
        assert sorted(map(key, ch)) == sorted(map(key, gh))
        results.append(sorted(map(key, gh)))
    print(results)
