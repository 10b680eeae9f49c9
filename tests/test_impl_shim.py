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
