"""GPU parity: Backward parser (a10) + domain decoding (a11) through the C ABI vs the CPU oracle.

Bars: Forward/Backward scores within 1e-3 nat of the oracle (north_star: 0.01 bit = 6.9e-3 nat) and of each other
(the reference's own fwd-vs-bck bar, src/impl_sse/fwdback_fs.c:3191); X rows within 1e-4 relative; btot/etot/mocc
within 1e-3 absolute (the reference's SIMD-vs-generic bar with exact logsum, src/impl_sse/decoding_fs.c:624-683).
"""
import ctypes as C

import numpy as np
import pytest

import common
from test_gpu_fs_forward import make_block

pytestmark = pytest.mark.gpu


def oracle_bck_decode(po, model, dsq, start, L, xf5_loop):
    lib = po.lib()
    lib.bo_fs_oprofile_ReconfigLength(model.om_fs3, L // 3)
    oxf, oxb = lib.bo_mx_create(model.M, L, 0), lib.bo_mx_create(model.M, L, 0)
    fsc, bsc = C.c_float(), C.c_float()
    sub = np.ascontiguousarray(dsq[start - 1: start + L + 1])
    st = lib.bo_ForwardParser_Frameshift_3Codons(po.u8ptr(sub), L, model.om_fs3, oxf, C.byref(fsc))
    if st == 0:
        st = lib.bo_BackwardParser_Frameshift_3Codons(po.u8ptr(sub), L, model.om_fs3, oxf, oxb, C.byref(bsc))
    out = {"status": st, "fwdsc": fsc.value, "bcksc": bsc.value}
    if st == 0:
        btot, etot, mocc = (np.zeros(L + 1, np.float32) for _ in range(3))
        x5 = np.asarray(xf5_loop, np.float32)
        assert lib.bo_DomainDecoding_Frameshift(po.fptr(x5), oxf, oxb, po.fptr(btot), po.fptr(etot), po.fptr(mocc)) == 0
        out.update(btot=btot, etot=etot, mocc=mocc, fx=po.mx_xmx(oxf).copy(), bx=po.mx_xmx(oxb).copy())
    lib.bo_mx_destroy(oxf)
    lib.bo_mx_destroy(oxb)
    return out


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("2OG-FeII_Oxy_3.bhmm", 0), ("tRNA-synthetases.bhmm", 1),
                                           ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0), ("MET-ct4.bhmm", 0), ("synthetic_M624.bhmm", 0), ("synthetic_M903.bhmm", 0)])
def test_backward_and_domain_decoding_match_oracle(oracle, gpu_ctx, hmmfile, index):
    po = oracle
    from bath_b200 import capi
    model = po.Model(common.golden(hmmfile), index)
    rng = np.random.default_rng(7 + index)
    dsq, wins = make_block(rng, model, n_random=10, n_homolog=22,
                           lengths=[60, 61, 62, 63, 300, 449, 600, 3 * max(model.max_length, 300) + 5])
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    gpu_ctx.upload_block(dsq)
    w = capi.Context.make_windows([s for s, _ in wins], [l for _, l in wins], nj=1.0)
    xf5 = (0.97, 0.97, 0.97)
    mocc, btot, etot, fsc, bsc, st = gpu_ctx.fs_bck_decode(w, (0.5, 0.5), xf5)
    fxs, bxs = gpu_ctx.fs_fetch_xrows(0, w), gpu_ctx.fs_fetch_xrows(1, w)
    worst = {"fwd": 0.0, "bck": 0.0, "mocc": 0.0, "btot": 0.0, "etot": 0.0, "rescaled_rows": 0}
    for t, (s, L) in enumerate(wins):
        o = oracle_bck_decode(po, model, dsq, s, L, xf5)
        assert st[t] == o["status"], (t, st[t], o["status"])
        if o["status"] != 0:
            continue
        assert abs(fsc[t] - o["fwdsc"]) <= 1e-3 and abs(bsc[t] - o["bcksc"]) <= 1e-3, (t, L, fsc[t], o["fwdsc"], bsc[t], o["bcksc"])
        assert abs(fsc[t] - bsc[t]) <= 1e-3
        # SCALE rows must be the same rows (same rescale trigger), values close
        assert np.array_equal(fxs[t][:, 5] > 1.0, o["fx"][:, 5] > 1.0), t
        assert np.array_equal(bxs[t][:, 5] > 1.0, o["bx"][:, 5] > 1.0), t
        np.testing.assert_allclose(fxs[t], o["fx"], rtol=2e-4, atol=1e-30)
        np.testing.assert_allclose(bxs[t], o["bx"], rtol=2e-4, atol=1e-30)
        for name, got in (("mocc", mocc[t]), ("btot", btot[t]), ("etot", etot[t])):
            d = float(np.max(np.abs(got - o[name])))
            worst[name] = max(worst[name], d)
            assert d <= 1e-3, (t, L, name, d)
        worst["fwd"] = max(worst["fwd"], abs(fsc[t] - o["fwdsc"]))
        worst["bck"] = max(worst["bck"], abs(bsc[t] - o["bcksc"]))
        worst["rescaled_rows"] += int((o["fx"][:, 5] > 1.0).sum())
    assert worst["rescaled_rows"] > 0, "the block must exercise the rescaling path"
    print(f"{hmmfile}[{index}] M={model.M}: {worst}")


def test_backward_rejects_short_windows(gpu_ctx, oracle):
    from bath_b200 import capi
    model = oracle.Model(common.golden("AMP_N.bhmm"))
    dsq = common.random_dna(np.random.default_rng(0), 100)
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    gpu_ctx.upload_block(dsq)
    w = capi.Context.make_windows([1], [4])
    with pytest.raises(capi.BathGpuError) as e:
        gpu_ctx.fs_bck_decode(w, (0.5, 0.5), (0.9, 0.9, 0.9))
    assert e.value.code == capi.EINVAL
