"""GPU parity: the per-envelope domain stage (a12-a15) through the C ABI vs the CPU oracle:
5-codon Forward/Backward scores, the posterior matrix, the optimal-accuracy matrix and score, the
traceback (states, nodes, positions and codon lengths identical; posteriors close) and null2.

Bars: scores within 1e-3 nat (north_star: 0.01 bit); posterior cells within 1e-4 absolute (the reference's own
SIMD-vs-generic bar is 1e-3 with exact logsum, src/impl_sse/decoding_fs.c:534,579); OA cells within 1e-3;
traces identical.
"""
import ctypes as C

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def oracle_domain(po, model, dsq, start, L):
    lib = po.lib()
    sub = np.ascontiguousarray(dsq[start - 1: start + L + 1])
    lib.bo_fs_oprofile_ReconfigUnihit(model.om_fs5, L // 3)
    fwd, bck, oa = lib.bo_mx_create(model.M, L, 8), lib.bo_mx_create(model.M, L, 3), lib.bo_mx_create(model.M, L, 3)
    fsc, bsc, e = C.c_float(), C.c_float(), C.c_float()
    out = {}
    st = lib.bo_Forward_Frameshift(po.u8ptr(sub), L, model.om_fs5, fwd, C.byref(fsc))
    if st == 0:
        st = lib.bo_Backward_Frameshift(po.u8ptr(sub), L, model.om_fs5, fwd, bck, C.byref(bsc))
    if st == 0:
        st = lib.bo_Decoding_Frameshift(model.om_fs5, fwd, bck)
    out.update(status=st, fwdsc=fsc.value, bcksc=bsc.value)
    if st == 0:
        assert lib.bo_OptimalAccuracy_Frameshift(model.om_fs5, fwd, oa, C.byref(e)) == 0
        tr = lib.bo_trace_create()
        assert lib.bo_OATrace_Frameshift(model.om_fs5, fwd, oa, tr) == 0
        out.update(oasc=e.value, trace=po.trace_list(tr), pp=po.mx_dp(fwd).copy(), ppx=po.mx_xmx(fwd).copy(),
                   oa=po.mx_dp(oa).copy(), oax=po.mx_xmx(oa).copy())
        null2 = np.zeros(29, np.float32)
        assert lib.bo_Null2_fs_ByExpectation(model.om_fs5, fwd, po.fptr(null2)) == 0
        out["null2"] = null2
        lib.bo_trace_destroy(tr)
    for mx in (fwd, bck, oa):
        lib.bo_mx_destroy(mx)
    lib.bo_fs_oprofile_ReconfigMultihit(model.om_fs5, 100)
    return out


def make_envelopes(rng, model, n_homolog, n_random):
    mat = common.hmm_mat(model)
    pieces, envs, pos = [], [], 1
    for t in range(n_homolog + n_random):
        if t < n_homolog:
            ins = common.sample_homolog(rng, mat, fs_rate=0.03, stop_rate=0.01)
            if t % 3 == 1:
                ins = ins[len(ins) // 4: 3 * len(ins) // 4]          # partial (local) homolog
            flank = int(rng.integers(0, 25))
            seg = np.concatenate([rng.integers(0, 4, flank), ins, rng.integers(0, 4, int(rng.integers(0, 25)))]).astype(np.uint8)
            if t % 5 == 2:
                seg[rng.integers(0, len(seg), 2)] = 15            # N's inside an envelope
        else:
            seg = rng.integers(0, 4, int(rng.integers(30, 200))).astype(np.uint8)
        gap = rng.integers(0, 4, int(rng.integers(0, 9))).astype(np.uint8)
        pieces += [gap, seg]
        pos += len(gap)
        envs.append((pos, len(seg)))
        pos += len(seg)
    body = np.concatenate(pieces)
    dsq = np.full(len(body) + 2, 255, np.uint8)
    dsq[1:-1] = body
    return dsq, envs


ST = {"M": 1, "D": 2, "I": 3, "S": 4, "N": 5, "B": 6, "E": 7, "C": 8, "T": 9, "J": 10}


def check_envelope(po, model, ctx, dsq, envs, res, tr, t, stats):
    s, L = envs[t]
    o = oracle_domain(po, model, dsq, s, L)
    assert res["status"][t] == o["status"], (t, res["status"][t], o["status"])
    if o["status"] != 0:
        return
    assert abs(res["envsc"][t] - o["fwdsc"]) <= 1e-3, (t, L, res["envsc"][t], o["fwdsc"])
    assert abs(res["bcksc"][t] - o["bcksc"]) <= 1e-3, (t, L, res["bcksc"][t], o["bcksc"])
    pp, oa, ppx, oax = ctx.fs_fetch_domain_matrices(t, model.M, L)
    dpp = float(np.max(np.abs(pp[:, :, 1:] - o["pp"][:, :, 1:])))
    assert dpp <= 1e-4, (t, L, "pp", dpp)
    assert float(np.max(np.abs(ppx[1:, [1, 2, 4]] - o["ppx"][1:, [1, 2, 4]]))) <= 1e-4
    fin = np.isfinite(o["oa"])
    assert np.array_equal(np.isfinite(oa[1:]), fin[1:]), (t, "OA -inf pattern")
    doa = float(np.max(np.abs(oa[1:][fin[1:]] - o["oa"][1:][fin[1:]]))) if fin[1:].any() else 0.0
    assert doa <= 1e-3, (t, L, "oa", doa)
    assert abs(res["oasc"][t] - o["oasc"]) <= 1e-3 * max(1.0, abs(o["oasc"]))
    np.testing.assert_allclose(res["null2"][t], o["null2"], rtol=1e-4, atol=1e-6)
    got = tr[res["trace_offset"][t]: res["trace_offset"][t] + res["trace_len"][t]]
    want = o["trace"]
    assert len(got) == len(want), (t, L, len(got), len(want))
    for z, (st_, k, i, c, p) in enumerate(want):
        g = got[z]
        assert (int(g["st"]), int(g["k"]), int(g["i"]), int(g["c"])) == (ST[st_], k, i, c), (t, z, g, want[z])
        assert abs(float(g["pp"]) - p) <= 1e-4
    stats["pp"] = max(stats["pp"], dpp)
    stats["oa"] = max(stats["oa"], doa)
    stats["fwd"] = max(stats["fwd"], abs(res["envsc"][t] - o["fwdsc"]))
    stats["steps"] += len(want)


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("2OG-FeII_Oxy_3.bhmm", 0), ("tRNA-synthetases.bhmm", 1),
                                           ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0), ("MET-ct4.bhmm", 0), ("synthetic_M624.bhmm", 0), ("synthetic_M903.bhmm", 0)])
def test_domain_stage_matches_oracle(oracle, gpu_ctx, hmmfile, index):
    po = oracle
    from bath_b200 import capi
    model = po.Model(common.golden(hmmfile), index)
    rng = np.random.default_rng(11 + index)
    dsq, envs = make_envelopes(rng, model, n_homolog=9, n_random=4)
    gpu_ctx.load_fs_profile(5, model.rfv(5), model.tfv(5))
    gpu_ctx.upload_block(dsq)
    e = capi.Context.make_windows([s for s, _ in envs], [l for _, l in envs], nj=0.0)     # unihit length model
    res, tr = gpu_ctx.fs_domains(e, xfE5=(1.0, 0.0))
    stats = {"pp": 0.0, "oa": 0.0, "fwd": 0.0, "steps": 0}
    for t in range(len(envs)):
        check_envelope(po, model, gpu_ctx, dsq, envs, res, tr, t, stats)
    print(f"{hmmfile}[{index}] M={model.M}: {len(envs)} envelopes, {stats}")


def test_amp_n_golden_alignment_on_gpu(oracle, gpu_ctx):
    """The hit of tutorial/AMP_N-fs.tbl: hmm 1-131, target 1-402, CIGAR, through the GPU domain stage."""
    po = oracle
    from bath_b200 import capi
    from test_oracle_golden import display
    model = po.Model(common.golden("AMP_N.bhmm"))
    _, _, seq = po.read_fasta(common.golden("target-AMP_N.fa"))[0]
    dsq = po.digitize_dna(seq)
    gpu_ctx.load_fs_profile(5, model.rfv(5), model.tfv(5))
    gpu_ctx.upload_block(dsq)
    e = capi.Context.make_windows([1], [len(seq)], nj=0.0)
    res, tr = gpu_ctx.fs_domains(e, xfE5=(1.0, 0.0))
    assert res["status"][0] == 0
    names = {v: k for k, v in ST.items()}
    trace = [(names[int(s["st"])], int(s["k"]), int(s["i"]), int(s["c"]), float(s["pp"])) for s in tr[: res["trace_len"][0]]]
    gm = model.gm_fs5.contents
    mc = gm.maxcodons
    codons = np.ctypeslib.as_array(gm.codons, shape=((model.M + 1) * (mc + 1),))[: (model.M + 1) * mc].reshape(model.M + 1, mc)
    indel = np.ctypeslib.as_array(gm.indel_pos, shape=((model.M + 1) * (mc + 1),))[: (model.M + 1) * mc].reshape(model.M + 1, mc)
    d = display(trace, dsq, codons, indel, model.hmm.contents.consensus.decode())
    tbl = [l for l in open(common.golden("AMP_N-fs.tbl")) if not l.startswith("#")][0].split()
    assert (d["hmm_from"], d["hmm_to"], d["ali_from"], d["ali_to"]) == (int(tbl[6]), int(tbl[7]), int(tbl[9]), int(tbl[10]))
    assert d["cigar"] == tbl[17] and d["shifts"] == int(tbl[15]) and d["stops"] == int(tbl[16])
