"""GPU: the multi-domain branch of the standard-translation domain definition (f2).
 * bathgpu_orf_forward_matrices (orf_forward_kernel<J, true, true> + export) against the oracle's p7_Forward matrix: every M / D / I
   cell within 2e-4 of the row maximum, X rows 2e-4 relative, scores 1e-3 nat;
 * the default pipeline on ORFs with abutting partial copies of the homolog: the GPU search resolves the flagged region exactly as the
   CPU-backend search does (byte-identical tables, same counters)."""
import ctypes as C

import numpy as np
import pytest

import common
from test_multidomain_std_cpu import STD_CASES, run_std_search

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("tRNA-synthetases.bhmm", 2), ("MET-ct4.bhmm", 0)])
def test_protein_forward_matrix_matches_oracle(oracle, gpu_ctx, hmmfile, index):
    from bath_b200 import capi
    po, L = oracle, oracle.lib()
    model = po.Model(common.golden(hmmfile), index)
    M = model.M
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    rng = np.random.default_rng(3 + index)
    mat = common.hmm_mat(model)
    p = mat[1:] / mat[1:].sum(axis=1, keepdims=True)
    hom = np.array([rng.choice(20, p=p[k]) for k in range(M)])
    res = np.concatenate([rng.integers(0, 20, 40), hom, hom[M // 3:], rng.integers(0, 20, 25), hom[: M // 2], rng.integers(0, 20, 30)]).astype(np.uint8)
    gpu_ctx.upload_orfs(res)
    n = len(res)
    regions = [(0, n), (20, min(n - 20, 2 * M)), (n // 2, n - n // 2 - 3), (5, 40)]
    regs = np.zeros(len(regions), capi.window_dtype)
    for r, (off, Lr) in enumerate(regions):
        regs[r]["start"], regs[r]["L"] = off, Lr
        pm = np.float32(3.0) / (np.float32(n) + np.float32(3.0))               # multihit at the ORF's length
        regs[r]["pmove"], regs[r]["ploop"] = pm, np.float32(1.0) - pm
    mxs, xrs, sc, st = gpu_ctx.orf_forward_matrices(regs, M)
    om = model.om
    L.bo_oprofile_ReconfigMultihit(om, n)
    for r, (off, Lr) in enumerate(regions):
        sub = np.concatenate([[255], res[off: off + Lr], [255]]).astype(np.uint8)
        fwd = L.bo_mx_create(M, Lr, 3)
        osc = C.c_float(0)
        ost = L.bo_Forward(po.u8ptr(sub), Lr, om, fwd, C.byref(osc))
        assert st[r] == ost
        if ost == 0:
            assert abs(sc[r] - osc.value) <= 1e-3
            dp = po.mx_dp(fwd)
            rowmax = np.maximum(dp.reshape(Lr + 1, -1).max(axis=1), 1e-30)[:, None, None]
            assert np.max(np.abs(mxs[r][:, :, :3] - dp) / rowmax) <= 2e-4
            assert (mxs[r][:, :, 3] == 0).all() and (mxs[r][:, 0, :] == 0).all()
            np.testing.assert_allclose(xrs[r], po.mx_xmx(fwd), rtol=2e-4, atol=1e-30)
        L.bo_mx_destroy(fwd)


def test_gpu_resolves_multidomain_regions_like_the_cpu_backend(oracle, gpu_ctx):
    for pieces, seed, ncopies in STD_CASES:
        be, keep = oracle.cpu_backend(4)
        chits, cst, ctbl, _, _ = run_std_search(pieces, backend=be, seed=seed)
        del keep
        ghits, gst, gtbl, _, _ = run_std_search(pieces, gpu_ctx=gpu_ctx, seed=seed)
        assert gtbl == ctbl
        for key in ("n_regions", "n_multidomain_regions", "n_envelopes", "n_hits_reported"):
            assert gst[key] == cst[key], key
        assert gst["n_multidomain_regions"] == 1 and len(ghits) == ncopies
