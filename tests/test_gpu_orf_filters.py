"""GPU parity: the integer ORF filters (a2-a4) through the C ABI vs the CPU oracle -- BIT-EXACT.

MSV score and status, Viterbi-filter score and status, the windows of p7_ViterbiFilter_BATH and p7_SSVFilter_BATH
(target position, model position, length, score) for both SSE and AVX2 stripe geometries.
"""
import ctypes as C

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu
AA = "ACDEFGHIKLMNPQRSTVWY"
LOG2 = 0.69314718055994529


def make_orfs(rng, model, n_random, n_homolog):
    """amino-acid ORFs: random ones and ones holding (pieces of) sequences emitted from the model's match states"""
    mat = common.hmm_mat(model)
    seqs = []
    for t in range(n_random + n_homolog):
        if t < n_random:
            s = rng.integers(0, 20, int(rng.integers(20, 420)))
            if t % 9 == 4:
                s[rng.integers(0, len(s), 2)] = 26                      # X residues
        else:
            p = mat[1:] / mat[1:].sum(axis=1, keepdims=True)
            hom = np.array([rng.choice(20, p=p[k]) for k in range(model.M)])
            mode = t % 4
            if mode == 1:
                hom = hom[len(hom) // 3:]
            elif mode == 2:
                hom = np.concatenate([hom[: len(hom) // 2], rng.integers(0, 20, 30), hom])   # two segments: the J state matters
            elif mode == 3:
                mask = rng.random(len(hom)) < 0.25
                hom[mask] = rng.integers(0, 20, int(mask.sum()))        # diverged
            s = np.concatenate([rng.integers(0, 20, int(rng.integers(0, 40))), hom, rng.integers(0, 20, int(rng.integers(0, 40)))])
        seqs.append(s.astype(np.uint8))
    return seqs


def host_params(po, model, lanes_u8, lanes_i16):
    o = model.om.contents
    return dict(M=model.M, tbm_b=o.tbm_b, tec_b=o.tec_b, base_b=o.base_b, bias_b=o.bias_b, scale_b=o.scale_b,
                base_w=o.base_w, ddbound_w=o.ddbound_w, xw_E_move=o.xw[0][0], xw_E_loop=o.xw[0][1], scale_w=o.scale_w,
                cpu_lanes_u8=lanes_u8, cpu_lanes_i16=lanes_i16)


def describe(po, model, seqs, P, filtersc_of):
    """bathgpu_orf descriptors with the host-owned per-length integers, computed by the oracle's own conversion code"""
    from bath_b200 import capi
    lib = po.lib()
    o = model.om.contents
    orfs = np.zeros(len(seqs), capi.orf_dtype)
    off = 0
    for t, s in enumerate(seqs):
        L = len(s)
        lib.bo_oprofile_ReconfigLength(model.om, L)
        orfs[t]["offset"], orfs[t]["L"] = off, L
        orfs[t]["tjb_b"], orfs[t]["xw_move"] = o.tjb_b, o.xw[1][0]
        fsc = filtersc_of(L)
        invP = np.float32(lib.bo_gumbel_invsurv(P, o.evparam[2], o.evparam[3]))
        orfs[t]["vit_thresh"] = np.int16(np.ceil((np.float64(np.float32(fsc)) + LOG2 * np.float64(invP) + 3.0) * np.float64(o.scale_w)
                                                 - float(o.xw[0][0]) - float(o.xw[3][0]) + float(o.base_w)))
        invP = np.float32(lib.bo_gumbel_invsurv(P, o.evparam[0], o.evparam[1]))
        ext = int(np.ceil((np.float64(np.float32(fsc)) + LOG2 * np.float64(invP) + 3.0) * np.float64(o.scale_b) + o.base_b + o.tec_b + o.tjb_b))
        orfs[t]["ext_thresh"] = ext
        orfs[t]["ssv_thresh"] = ext & 0xff
        orfs[t]["flags"] = 1
        off += L
    return orfs


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("2OG-FeII_Oxy_3.bhmm", 0), ("tRNA-synthetases.bhmm", 1),
                                           ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0), ("MET-ct4.bhmm", 0), ("synthetic_M624.bhmm", 0), ("synthetic_M903.bhmm", 0)])
@pytest.mark.parametrize("lanes", [(16, 8), (32, 16)])
def test_filters_bit_exact(oracle, gpu_ctx, hmmfile, index, lanes):
    po = oracle
    lib = po.lib()
    model = po.Model(common.golden(hmmfile), index)
    rng = np.random.default_rng(5 + index)
    seqs = make_orfs(rng, model, n_random=40, n_homolog=24)
    rbv, rwv, twv, _, _ = model.om_tables()
    gpu_ctx.load_filter_profile(host_params(po, model, *lanes), rbv, rwv, twv)
    gpu_ctx.upload_orfs(np.concatenate(seqs))
    P = 1e-3                                                  # pli->F2 (src/p7_pipeline.c:168)
    bg = model.bg

    def nullsc(L):
        lib.bo_bg_SetLength(bg, L)
        return lib.bo_bg_NullOne(bg, L)

    orfs = describe(po, model, seqs, P, nullsc)
    msv, mst = gpu_ctx.msv_orfs(orfs)
    vit, vst, vwin = gpu_ctx.vit_orfs(orfs)
    swin = gpu_ctx.ssv_windows(orfs)
    ssv = model.ssv_scores()
    wl = po.WINDOWLIST()
    n_overflow = n_vwin = n_swin = 0
    for t, s in enumerate(seqs):
        L = len(s)
        d = np.concatenate([[255], s, [255]]).astype(np.uint8)
        lib.bo_oprofile_ReconfigLength(model.om, L)
        a, v = C.c_float(), C.c_float()
        st = lib.bo_MSVFilter(po.u8ptr(d), L, model.om, C.byref(a))
        assert mst[t] == st, (t, L, mst[t], st)
        assert msv[t] == a.value or (np.isinf(msv[t]) and np.isinf(a.value)), (t, L, msv[t], a.value)
        n_overflow += st == 16
        lib.bo_windowlist_reset(C.byref(wl))
        fsc = nullsc(L)
        st = lib.bo_ViterbiFilter_BATH(po.u8ptr(d), L, model.om, po.u8ptr(ssv), fsc, P, lanes[1], C.byref(wl), C.byref(v))
        assert vst[t] == st and (vit[t] == v.value or (np.isinf(vit[t]) and np.isinf(v.value))), (t, L, vit[t], v.value, vst[t], st)
        want = [(n, k, ln) for (n, k, ln, sc) in po.windows(wl)] if st == 0 else None
        got = [(int(w["n"]), int(w["k"]), int(w["length"])) for w in vwin[vwin["orf"] == t]]
        if want is not None:
            assert got == want, (t, L, got, want)
            n_vwin += len(want)
        lib.bo_windowlist_reset(C.byref(wl))
        assert lib.bo_SSVFilter_BATH(po.u8ptr(d), L, model.om, po.u8ptr(ssv), fsc, P, lanes[0], C.byref(wl)) == 0
        want = po.windows(wl)
        got = [(int(w["n"]), int(w["k"]), int(w["length"]), float(w["score"])) for w in swin[swin["orf"] == t]]
        assert got == [(n, k, ln, np.float32(sc)) for (n, k, ln, sc) in want], (t, L, got, want)
        n_swin += len(want)
    lib.bo_windowlist_free(C.byref(wl))
    print(f"{hmmfile}[{index}] lanes={lanes}: {n_vwin} Viterbi windows, {n_swin} SSV windows, {n_overflow} MSV overflows")
    # (whole homologs of the long models all overflow the byte filter, which leaves no Viterbi-window cases for them)
    assert n_swin > 0 and n_overflow > 0 and (n_vwin > 0 or model.M > 400)
    print(f"{hmmfile}[{index}] lanes={lanes}: {len(seqs)} ORFs, {n_overflow} MSV overflows, {n_vwin} Viterbi windows, {n_swin} SSV windows")


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("tRNA-synthetases.bhmm", 1), ("PTHR37536.bhmm", 0), ("MET-ct4.bhmm", 0), ("synthetic_M624.bhmm", 0), ("synthetic_M903.bhmm", 0)])
def test_protein_forward_parser(oracle, gpu_ctx, hmmfile, index):
    """a6: p7_ForwardParser over ORFs, within 1e-3 nat of the oracle; statuses equal"""
    po = oracle
    lib = po.lib()
    model = po.Model(common.golden(hmmfile), index)
    rng = np.random.default_rng(9 + index)
    seqs = make_orfs(rng, model, n_random=30, n_homolog=30)
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    gpu_ctx.upload_orfs(np.concatenate(seqs))
    from bath_b200 import capi
    orfs = np.zeros(len(seqs), capi.orf_dtype)
    off = 0
    for t, s in enumerate(seqs):
        orfs[t]["offset"], orfs[t]["L"] = off, len(s)
        off += len(s)
    sc, st = gpu_ctx.fwd_orfs(orfs, nj=1.0, xfE=(0.5, 0.5))
    worst = 0.0
    for t, s in enumerate(seqs):
        d = np.concatenate([[255], s, [255]]).astype(np.uint8)
        lib.bo_oprofile_ReconfigLength(model.om, len(s))
        v = C.c_float()
        ost = lib.bo_ForwardParser(po.u8ptr(d), len(s), model.om, C.byref(v))
        assert st[t] == ost
        if ost == 0:
            worst = max(worst, abs(sc[t] - v.value))
            assert abs(sc[t] - v.value) <= 1e-3, (t, len(s), sc[t], v.value)
    print(f"{hmmfile}[{index}]: {len(seqs)} ORFs, max |dsc| = {worst:.2e} nat, max score {sc.max():.1f}")
