"""The drop-in boundary made checkable: integration/impl_cuda_shim.c implements the reference's impl-layer prototypes
(src/impl_sse/impl_sse.h:493-494, verbatim signatures) on libbathgpu.so, compiled against a stand-in for impl_sse.h that mirrors
the field layout of P7_FS_OPROFILE / P7_OMX (src/impl_sse/impl_sse.h:200-244, :329-358).

CPU: the library builds, exports the reference's symbols, and un-striping inverts the SSE striping of fs_fb_conversion
(src/impl_sse/p7_fs_oprofile.c:222-296) on real profiles.  GPU: p7_ForwardParser_Frameshift_3Codons /
p7_BackwardParser_Frameshift_3Codons called the way p7_pli_Frameshift calls them (src/p7_pipeline.c:1450, :1469-1470) return the
oracle's scores, X rows, totscale and status.
"""
import ctypes as C

import numpy as np
import pytest

import common


def shim():
    from bath_b200 import build
    build.build_library()
    L = C.CDLL(build.build_shim_library())
    fp, vp = C.POINTER(C.c_float), C.c_void_p
    L.shimtest_make_oprofile.restype = vp
    L.shimtest_make_oprofile.argtypes = [C.c_int, C.c_int, C.c_int, fp, fp, fp]
    L.shimtest_free_oprofile.argtypes = [vp]
    L.shimtest_make_omx.restype = vp
    L.shimtest_make_omx.argtypes = [C.c_int]
    L.shimtest_free_omx.argtypes = [vp]
    L.shimtest_omx_xmx.restype = fp
    L.shimtest_omx_xmx.argtypes = [vp]
    L.shimtest_omx_totscale.restype = C.c_float
    L.shimtest_omx_totscale.argtypes = [vp]
    L.shimtest_omx_field.argtypes = [vp, C.c_int]
    L.shimtest_set_length.argtypes = [vp, C.c_float, C.c_float]
    L.bathshim_unstripe_fs_profile.argtypes = [vp, fp, fp]
    L.p7_ForwardParser_Frameshift_3Codons.argtypes = [C.POINTER(C.c_uint8), C.c_int, vp, vp, vp, fp]
    L.p7_BackwardParser_Frameshift_3Codons.argtypes = [C.POINTER(C.c_uint8), C.c_int, vp, vp, vp, vp, fp]
    return L


def f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def striped(L, model, which):
    from bath_b200 import hostapi
    rfv = np.ascontiguousarray(model.rfv(which), np.float32)
    tfv = np.ascontiguousarray(model.tfv(which), np.float32)
    xf = np.array([0.5, 0.5, 0.03, 0.97, 0.03, 0.97, 0.03, 0.97], np.float32)          # E, N, J, C x {MOVE, LOOP}
    om = L.shimtest_make_oprofile(which, model.M, rfv.shape[0], f32p(rfv), f32p(tfv), f32p(xf))
    return om, rfv, tfv


def test_shim_exports_the_reference_symbols():
    L = shim()
    for name in ("p7_ForwardParser_Frameshift_3Codons", "p7_BackwardParser_Frameshift_3Codons", "bathshim_unstripe_fs_profile",
                 "bathshim_set_device", "bathshim_release"):
        assert hasattr(L, name), name


@pytest.mark.parametrize("hmmfile,index,which", [("AMP_N.bhmm", 0, 3), ("AMP_N.bhmm", 0, 5), ("tRNA-synthetases.bhmm", 1, 3),
                                                 ("PTHR37536.bhmm", 0, 3), ("MET-ct4.bhmm", 0, 5)])
def test_unstriping_inverts_the_sse_striping(hmmfile, index, which):
    from bath_b200 import hostapi
    L = shim()
    model = hostapi.QueryModel(common.golden(hmmfile), index)
    om, rfv, tfv = striped(L, model, which)
    M = model.M
    r2, t2 = np.full_like(rfv, -1.0), np.full_like(tfv, -1.0)
    assert L.bathshim_unstripe_fs_profile(om, f32p(r2), f32p(t2)) == rfv.shape[0]
    L.shimtest_free_oprofile(om)
    assert np.array_equal(r2[:, 1:], rfv[:, 1:])                       # every emission row, nodes 1..M, bit for bit
    assert (r2[:, 0] == 0).all()
    # transitions: BM MM IM DM out of nodes 0..M-1; MD MI II DD out of nodes 1..M-1 (the striped form holds nothing for node M:
    # p7_fs_oprofile.c:271, :281 -- the kernels never read those entries)
    assert np.array_equal(t2[:4, :M], tfv[:4, :M])
    assert np.array_equal(t2[4:, 1:M], tfv[4:, 1:M])
    assert (t2[4:, M] == 0).all() and (t2[4:, 0] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("tRNA-synthetases.bhmm", 2)])
def test_parsers_through_the_reference_signatures_match_the_oracle(oracle, hmmfile, index):
    from bath_b200 import hostapi
    from test_gpu_fs_forward import make_block
    from test_gpu_fs_backward import oracle_bck_decode
    po = oracle
    L = shim()
    hmodel = hostapi.QueryModel(common.golden(hmmfile), index)
    omodel = po.Model(common.golden(hmmfile), index)
    om, _, _ = striped(L, hmodel, 3)
    rng = np.random.default_rng(31 + index)
    dsq, wins = make_block(rng, omodel, n_random=3, n_homolog=5, lengths=[60, 301, 449, 3 * max(omodel.max_length, 300) + 5])
    for s, Lw in wins:
        sub = np.ascontiguousarray(dsq[s - 1: s + Lw + 1])
        pm = np.float32(3.0) / (np.float32(Lw // 3) + np.float32(3.0))            # p7_fs_oprofile_ReconfigLength(om_fs3, L/3), nj = 1
        L.shimtest_set_length(om, pm, np.float32(1.0) - pm)
        oxf, oxb = L.shimtest_make_omx(Lw), L.shimtest_make_omx(Lw)
        fsc, bsc = C.c_float(), C.c_float()
        stf = L.p7_ForwardParser_Frameshift_3Codons(sub.ctypes.data_as(C.POINTER(C.c_uint8)), Lw, om, oxf, None, C.byref(fsc))
        stb = L.p7_BackwardParser_Frameshift_3Codons(sub.ctypes.data_as(C.POINTER(C.c_uint8)), Lw, om, oxf, oxb, None, C.byref(bsc))
        o = oracle_bck_decode(po, omodel, dsq, s, Lw, (0.97, 0.97, 0.97))
        assert stf == o["status"] and stb == o["status"], (Lw, stf, stb, o["status"])
        if o["status"] == 0:
            assert abs(fsc.value - o["fwdsc"]) <= 1e-3 and abs(bsc.value - o["bcksc"]) <= 1e-3, (Lw, fsc.value, o["fwdsc"], bsc.value, o["bcksc"])
            fx = np.ctypeslib.as_array(L.shimtest_omx_xmx(oxf), shape=(Lw + 1, 6))
            bx = np.ctypeslib.as_array(L.shimtest_omx_xmx(oxb), shape=(Lw + 1, 6))
            np.testing.assert_allclose(fx, o["fx"], rtol=2e-4, atol=1e-30)
            np.testing.assert_allclose(bx, o["bx"], rtol=2e-4, atol=1e-30)
            want_tot = float(np.sum(np.log(o["fx"][:, 5][o["fx"][:, 5] != 1.0].astype(np.float64))))
            assert abs(L.shimtest_omx_totscale(oxf) - want_tot) <= 1e-3 * max(1.0, abs(want_tot))
            assert (L.shimtest_omx_field(oxf, 0), L.shimtest_omx_field(oxf, 1), L.shimtest_omx_field(oxf, 2)) == (hmodel.M, Lw, 1)
            assert L.shimtest_omx_field(oxb, 2) == 0
        L.shimtest_free_omx(oxf)
        L.shimtest_free_omx(oxb)
    L.shimtest_free_oprofile(om)
    L.bathshim_release()


# ---- the protein profile: un-striping of P7_OPROFILE (mf_conversion / vf_conversion / fb_conversion layouts) and the three entry points

def striped_protein(L, po, model, lanes=(16, 8)):
    """P7_OPROFILE striped from the oracle's plain tables, with its scalar parameters"""
    rbv, rwv, twv, rfv, tfv = (np.ascontiguousarray(a) for a in model.om_tables())
    o = model.om.contents
    iprm = np.array([o.tbm_b, o.tec_b, o.tjb_b, o.base_b, o.bias_b, o.base_w, o.ddbound_w, o.xw[0][0], o.xw[0][1], o.xw[1][0]], np.int32)
    fprm = np.array([o.scale_b, o.scale_w, 1.0, 0.5, 0.5], np.float32)
    L.shimtest_make_oprofile_protein.restype = C.c_void_p
    L.shimtest_make_oprofile_protein.argtypes = [C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_int16), C.POINTER(C.c_float),
                                                 C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    om = L.shimtest_make_oprofile_protein(model.M, rbv.ctypes.data_as(C.POINTER(C.c_uint8)), rwv.ctypes.data_as(C.POINTER(C.c_int16)),
                                          twv.ctypes.data_as(C.POINTER(C.c_int16)), f32p(rfv), f32p(tfv),
                                          iprm.ctypes.data_as(C.POINTER(C.c_int32)), f32p(fprm))
    return om, (rbv, rwv, twv, rfv, tfv)


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("PTH2.bhmm", 0), ("tRNA-synthetases.bhmm", 2), ("MET-ct4.bhmm", 1)])
def test_unstriping_inverts_the_protein_profile_striping(oracle, hmmfile, index):
    po = oracle
    L = shim()
    model = po.Model(common.golden(hmmfile), index)
    M = model.M
    om, (rbv, rwv, twv, rfv, tfv) = striped_protein(L, po, model)
    b2, w2, t2 = np.zeros_like(rbv), np.zeros_like(rwv), np.zeros_like(twv)
    r2, f2 = np.zeros_like(rfv), np.zeros_like(tfv)
    L.bathshim_unstripe_oprofile.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_int16), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.bathshim_unstripe_oprofile(om, b2.ctypes.data_as(C.POINTER(C.c_uint8)), w2.ctypes.data_as(C.POINTER(C.c_int16)),
                                 t2.ctypes.data_as(C.POINTER(C.c_int16)), f32p(r2), f32p(f2))
    L.shimtest_free_oprofile_protein.argtypes = [C.c_void_p]
    L.shimtest_free_oprofile_protein(om)
    assert np.array_equal(b2[:, 1:], rbv[:, 1:]) and np.array_equal(w2[:, 1:], rwv[:, 1:]) and np.array_equal(r2[:, 1:], rfv[:, 1:])
    assert np.array_equal(t2[:4, :M], twv[:4, :M]) and np.array_equal(t2[4:, 1:M], twv[4:, 1:M])       # as for the frameshift profile: nothing for node M
    assert np.array_equal(f2[:4, :M], tfv[:4, :M]) and np.array_equal(f2[4:, 1:M], tfv[4:, 1:M])


@pytest.mark.gpu
@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("tRNA-synthetases.bhmm", 2)])
def test_filters_through_the_reference_signatures_match_the_oracle(oracle, hmmfile, index):
    """p7_MSVFilter / p7_ViterbiFilter bit-exact, p7_ForwardParser within 1e-3 nat, called as p7_Pipeline_BATH calls them
    (src/p7_pipeline.c:1649, :1672, :1779) after p7_oprofile_ReconfigLength(om, L)"""
    from test_gpu_orf_filters import make_orfs
    po, lib = oracle, oracle.lib()
    L = shim()
    model = po.Model(common.golden(hmmfile), index)
    om, _ = striped_protein(L, po, model)
    u8p, fp = C.POINTER(C.c_uint8), C.POINTER(C.c_float)
    for f in (L.p7_MSVFilter, L.p7_ViterbiFilter, L.p7_ForwardParser):
        f.argtypes = [u8p, C.c_int, C.c_void_p, C.c_void_p, fp]
    L.shimtest_set_orf_length.argtypes = [C.c_void_p, C.c_int, C.c_int]
    rng = np.random.default_rng(9 + index)
    n_inf = 0
    for s in make_orfs(rng, model, 10, 10):
        Ls = len(s)
        d = np.concatenate([[255], s, [255]]).astype(np.uint8)
        lib.bo_oprofile_ReconfigLength(model.om, Ls)
        o = model.om.contents
        L.shimtest_set_orf_length(om, int(o.tjb_b), int(o.xw[1][0]))
        got, want = C.c_float(), C.c_float()
        for shim_fn, ora_fn, exact in ((L.p7_MSVFilter, lib.bo_MSVFilter, True), (L.p7_ViterbiFilter, lib.bo_ViterbiFilter, True)):
            st = shim_fn(d.ctypes.data_as(u8p), Ls, om, None, C.byref(got))
            ost = ora_fn(po.u8ptr(d), Ls, model.om, C.byref(want))
            assert st == ost, (shim_fn, Ls, st, ost)
            assert got.value == want.value or (np.isinf(got.value) and np.isinf(want.value)), (Ls, got.value, want.value)
            n_inf += bool(np.isinf(want.value))
        st = L.p7_ForwardParser(d.ctypes.data_as(u8p), Ls, om, None, C.byref(got))
        ox = lib.bo_mx_create(model.M, Ls, 0)
        ost = lib.bo_Forward(po.u8ptr(d), Ls, model.om, ox, C.byref(want))
        lib.bo_mx_destroy(ox)
        assert st == ost and (ost != 0 or abs(got.value - want.value) <= 1e-3), (Ls, got.value, want.value)
    L.shimtest_free_oprofile_protein.argtypes = [C.c_void_p]
    L.shimtest_free_oprofile_protein(om)
    L.bathshim_release()
