"""GPU parity: the standard-translation branch through the C ABI vs the CPU oracle (oracle/orf_domain.c):
protein Forward / Backward parsers over ORFs with their X rows (what p7_DomainDecoding reads), and the per-envelope
stage -- Forward, Backward, posterior decoding, optimal accuracy, traceback, null2 -- plus the whole default pipeline
(bathsearch without --fs) against tutorial/PTH2.tbl and tutorial/AMP_N.out on the device.

Bars: scores within 1e-3 nat (north_star: 0.01 bit); X rows 2e-4 relative; posterior cells 1e-4 absolute; optimal-accuracy
cells 1e-3 + 5e-5 L; traces identical (state, node, residue); null2 1e-4 relative.
"""
import ctypes as C

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu

ST = {"M": 1, "D": 2, "I": 3, "S": 4, "N": 5, "B": 6, "E": 7, "C": 8, "T": 9, "J": 10}


def make_orfs(rng, model, n_homolog, n_random):
    """amino-acid sequences: samples from the model's match states (whole, partial, two copies in a row, with X residues)
    inside random flanks, and random sequences; returned with the concatenated residue buffer"""
    mat = common.hmm_mat(model)
    M = mat.shape[0] - 1
    seqs = []
    for t in range(n_homolog + n_random):
        if t < n_homolog:
            core = np.array([rng.choice(20, p=mat[k] / mat[k].sum()) for k in range(1, M + 1)], np.uint8)
            if t % 3 == 1:
                core = core[M // 4: 3 * M // 4]
            if t % 4 == 2:
                core = np.delete(core, slice(M // 3, M // 3 + 7))                 # a deletion
            if t % 4 == 3:
                core = np.insert(core, M // 2, rng.integers(0, 20, 5))            # an insertion
            if t % 5 == 4:
                core = np.concatenate([core, rng.integers(0, 20, 12).astype(np.uint8), core[: M // 2]])   # two domains
            s = np.concatenate([rng.integers(0, 20, int(rng.integers(0, 30))), core, rng.integers(0, 20, int(rng.integers(0, 30)))])
            if t % 6 == 5:
                s[rng.integers(0, len(s), 2)] = 26                                # X from a degenerate codon
        else:
            s = rng.integers(0, 20, int(rng.integers(20, 300)))
        seqs.append(s.astype(np.uint8))
    return seqs


def orf_descs(capi, seqs):
    orfs = np.zeros(len(seqs), capi.orf_dtype)
    off = 0
    for t, s in enumerate(seqs):
        orfs[t]["offset"], orfs[t]["L"] = off, len(s)
        off += len(s)
    return orfs


def dsq_of(s):
    return np.concatenate([[255], s, [255]]).astype(np.uint8)


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("2OG-FeII_Oxy_3.bhmm", 0), ("tRNA-synthetases.bhmm", 1), ("PTHR37536.bhmm", 0), ("MET-ct4.bhmm", 0), ("synthetic_M624.bhmm", 0), ("synthetic_M903.bhmm", 0)])
def test_orf_parsers_xrows_match_oracle(oracle, gpu_ctx, hmmfile, index):
    po, lib = oracle, oracle.lib()
    from bath_b200 import capi
    model = po.Model(common.golden(hmmfile), index)
    rng = np.random.default_rng(5 + index)
    seqs = make_orfs(rng, model, 10, 5)
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    gpu_ctx.upload_orfs(np.concatenate(seqs))
    fx, bx, fsc, bsc, st = gpu_ctx.orf_fwd_bck_xrows(orf_descs(capi, seqs), nj=1.0, xfE=(0.5, 0.5))
    worst = 0.0
    for t, s in enumerate(seqs):
        L, d = len(s), dsq_of(s)
        lib.bo_oprofile_ReconfigMultihit(model.om, L)
        oxf, oxb = lib.bo_mx_create(model.M, L, 0), lib.bo_mx_create(model.M, L, 0)
        f, b = C.c_float(), C.c_float()
        assert lib.bo_Forward(po.u8ptr(d), L, model.om, oxf, C.byref(f)) == st[t] == 0
        assert lib.bo_Backward(po.u8ptr(d), L, model.om, oxf, oxb, C.byref(b)) == 0
        assert abs(fsc[t] - f.value) <= 1e-3 and abs(bsc[t] - b.value) <= 1e-3, (t, L, fsc[t], f.value, bsc[t], b.value)
        of, ob = po.mx_xmx(oxf), po.mx_xmx(oxb)
        assert np.array_equal(fx[t][:, 5] > 1.0, of[:, 5] > 1.0)                  # the same rows rescale
        for got, want in ((fx[t], of), (bx[t], ob)):
            scale = np.maximum(np.abs(want), 1e-30)
            err = float(np.max(np.abs(got - want) / np.maximum(scale, np.max(np.abs(want), axis=0) * 1e-3)))
            worst = max(worst, err)
            assert err <= 2e-4, (t, L, err)
        lib.bo_mx_destroy(oxf); lib.bo_mx_destroy(oxb)
    print(f"{hmmfile}[{index}] M={model.M}: {len(seqs)} ORFs, worst relative X-row error {worst:.2e}")


def oracle_envelope(po, model, s):
    lib = po.lib()
    L, d = len(s), dsq_of(s)
    lib.bo_oprofile_ReconfigUnihit(model.om, L)
    fwd, bck = lib.bo_mx_create(model.M, L, 3), lib.bo_mx_create(model.M, L, 3)
    f, b, e = C.c_float(), C.c_float(), C.c_float()
    st = lib.bo_Forward(po.u8ptr(d), L, model.om, fwd, C.byref(f))
    if st == 0:
        st = lib.bo_Backward(po.u8ptr(d), L, model.om, fwd, bck, C.byref(b))
    if st == 0:
        st = lib.bo_Decoding(model.om, fwd, bck, bck)
    out = dict(status=st, fwdsc=f.value, bcksc=b.value)
    if st == 0:
        out.update(pp=po.mx_dp(bck).copy(), ppx=po.mx_xmx(bck).copy())
        assert lib.bo_OptimalAccuracy(model.om, bck, fwd, C.byref(e)) == 0
        tr = lib.bo_trace_create()
        assert lib.bo_OATrace(model.om, bck, fwd, 4, tr) == 0
        null2 = np.zeros(29, np.float32)
        assert lib.bo_Null2_ByExpectation(model.om, bck, po.fptr(null2)) == 0
        out.update(oasc=e.value, trace=po.trace_list(tr), oa=po.mx_dp(fwd).copy(), oax=po.mx_xmx(fwd).copy(), null2=null2)
        lib.bo_trace_destroy(tr)
    lib.bo_mx_destroy(fwd); lib.bo_mx_destroy(bck)
    lib.bo_oprofile_ReconfigMultihit(model.om, 100)
    return out


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("2OG-FeII_Oxy_3.bhmm", 0), ("tRNA-synthetases.bhmm", 1),
                                           ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0), ("MET-ct4.bhmm", 0), ("synthetic_M624.bhmm", 0), ("synthetic_M903.bhmm", 0)])
def test_orf_domain_stage_matches_oracle(oracle, gpu_ctx, hmmfile, index):
    po = oracle
    from bath_b200 import capi
    model = po.Model(common.golden(hmmfile), index)
    rng = np.random.default_rng(23 + index)
    seqs = make_orfs(rng, model, 12, 4)
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    gpu_ctx.upload_orfs(np.concatenate(seqs))
    offs = np.concatenate([[0], np.cumsum([len(s) for s in seqs])])
    envs = np.zeros(len(seqs), capi.window_dtype)
    for t, s in enumerate(seqs):
        envs[t]["start"], envs[t]["L"] = offs[t], len(s)
        envs[t]["pmove"] = np.float32(2.0) / (np.float32(len(s)) + np.float32(2.0))
        envs[t]["ploop"] = np.float32(1.0) - envs[t]["pmove"]
    res, tr = gpu_ctx.orf_domains(envs, xfE=(1.0, 0.0), M=model.M)
    stats = dict(pp=0.0, oa=0.0, fwd=0.0, steps=0)
    for t, s in enumerate(seqs):
        o, L = oracle_envelope(po, model, s), len(s)
        if o["status"] != 0 or abs(o["fwdsc"] - o["bcksc"]) > 1e-2:
            # two whole copies of a long model inside one unihit envelope: Backward underflows under Forward's scale factors in
            # the reference's own arithmetic (its Forward and Backward scores disagree or it returns eslERANGE); nothing to compare
            assert res["status"][t] in (0, 16)
            continue
        assert res["status"][t] == o["status"] == 0
        assert abs(res["envsc"][t] - o["fwdsc"]) <= 1e-3 and abs(res["bcksc"][t] - o["bcksc"]) <= 1e-3, (t, L, res["envsc"][t], o["fwdsc"])
        pp, oa, ppx, oax = gpu_ctx.orf_fetch_domain_matrices(t, L, model.M)
        dpp = float(np.max(np.abs(pp[1:] - o["pp"][1:])))
        assert dpp <= 1e-4, (t, L, "pp", dpp)
        assert float(np.max(np.abs(ppx[1:, [1, 2, 4]] - o["ppx"][1:, [1, 2, 4]]))) <= 1e-4
        fin = np.isfinite(o["oa"][1:])
        assert np.array_equal(np.isfinite(oa[1:]), fin), (t, "OA -inf pattern")
        doa = float(np.max(np.abs(oa[1:][fin] - o["oa"][1:][fin])))
        assert doa <= 1e-3 + 5e-5 * L, (t, L, "oa", doa)       # OA cells are running sums of up to L posteriors
        assert abs(res["oasc"][t] - o["oasc"]) <= 1e-3 * max(1.0, abs(o["oasc"]))
        np.testing.assert_allclose(res["null2"][t], o["null2"], rtol=1e-4, atol=1e-6)
        got = tr[res["trace_offset"][t]: res["trace_offset"][t] + res["trace_len"][t]]
        want = o["trace"]
        assert len(got) == len(want), (t, L, len(got), len(want))
        for z, (st_, k, i, c, p) in enumerate(want):
            g = got[z]
            assert (int(g["st"]), int(g["k"]), int(g["i"]), int(g["c"])) == (ST[st_], k, i, c), (t, z, g, want[z])
            assert abs(float(g["pp"]) - p) <= 1e-4
        stats["pp"] = max(stats["pp"], dpp); stats["oa"] = max(stats["oa"], doa)
        stats["fwd"] = max(stats["fwd"], abs(res["envsc"][t] - o["fwdsc"])); stats["steps"] += len(want)
    print(f"{hmmfile}[{index}] M={model.M}: {len(seqs)} envelopes, {stats}")


def test_orf_stage_argument_errors(oracle, gpu_ctx):
    from bath_b200 import capi
    model = oracle.Model(common.golden("AMP_N.bhmm"))
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    gpu_ctx.upload_orfs(np.arange(50, dtype=np.uint8) % 20)
    bad = np.zeros(1, capi.orf_dtype); bad["offset"], bad["L"] = 40, 20
    with pytest.raises(capi.BathGpuError) as e:
        gpu_ctx.orf_fwd_bck_xrows(bad)
    assert e.value.code == capi.EINVAL
    env = np.zeros(1, capi.window_dtype); env["start"], env["L"], env["pmove"], env["ploop"] = 0, 50, 0.04, 0.96
    with pytest.raises(capi.BathGpuError) as e:
        gpu_ctx.orf_domains(env, max_steps=3)
    assert e.value.code == capi.EINVAL and "trace" in str(e.value)


def test_default_pipeline_matches_golden_tables_on_gpu(gpu_ctx):
    """bathsearch without --fs on the device stages: tutorial/PTH2.tbl (4 hits, CIGARs, PID) + PTH2.out footer, tutorial/AMP_N.out"""
    from test_gpu_pipeline import run_search
    from test_host_pipeline_cpu import check_pth2_default, check_amp_n_default
    check_pth2_default(*run_search(gpu_ctx, "PTH2.bhmm", "target-PTH2.fa", std_only=1))
    check_amp_n_default(*run_search(gpu_ctx, "AMP_N.bhmm", "target-AMP_N.fa", std_only=1))


def test_fs_pipeline_standard_branch_on_gpu(gpu_ctx):
    """--fs on PTH2: hit 1 of tutorial/PTH2.tbl comes out of the standard-translation branch unchanged"""
    from test_gpu_pipeline import run_search
    from test_host_pipeline_cpu import _row
    hits, st = run_search(gpu_ctx, "PTH2.bhmm", "target-PTH2.fa")
    f = [l.split() for l in open(common.golden("PTH2.tbl")) if not l.startswith("#")][0]
    assert st["n_std_windows"] == 1 and len(hits) == 4
    assert _row(hits[0]) == (int(f[6]), int(f[7]), int(f[9]), int(f[10]), f"{float(f[11]):.2g}", f[12], f[13]) and hits[0]["cigar"] == f[15]


def test_default_pipeline_long_models_on_gpu(gpu_ctx):
    """tutorial/MET-ct4.out on the device: M = 409 and 458 run on the 16-nodes-per-lane instantiations (part of the row state
    in local memory), codon table 4, both strands"""
    from bath_b200 import hostapi
    from test_host_pipeline_cpu import check_met_ct4
    check_met_ct4(lambda model: hostapi.Search(model, gpu_ctx, std_only=1))
