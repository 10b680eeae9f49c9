"""GPU edge cases through the C ABI: argument errors (Easel status codes, never a crash or a fallback), shortest and longest
inputs, all-degenerate targets, single-residue ORFs -- each compared with the oracle where a result is defined."""
import ctypes as C

import numpy as np
import pytest

import common
from test_gpu_fs_forward import oracle_fwd

pytestmark = pytest.mark.gpu


def test_argument_errors(oracle, gpu_ctx):
    from bath_b200 import capi
    model = oracle.Model(common.golden("AMP_N.bhmm"))
    fresh = capi.Context(0)
    dsq = common.random_dna(np.random.default_rng(0), 500)
    w = capi.Context.make_windows([1], [300])
    with pytest.raises(capi.BathGpuError) as e:            # nothing uploaded / loaded yet
        fresh.fs_fwd_windows(w)
    assert e.value.code == capi.EINVAL
    fresh.load_fs_profile(3, model.rfv(3), model.tfv(3))
    with pytest.raises(capi.BathGpuError) as e:            # profile but no block
        fresh.fs_fwd_windows(w)
    assert e.value.code == capi.EINVAL and "block" in str(e.value)
    fresh.upload_block(dsq)
    for starts, lens in (([0], [100]), ([450], [100]), ([1], [2]), ([1], [501])):      # outside the block, too short
        with pytest.raises(capi.BathGpuError) as e:
            fresh.fs_fwd_windows(capi.Context.make_windows(starts, lens))
        assert e.value.code == capi.EINVAL
    with pytest.raises(capi.BathGpuError) as e:            # wrong table height for 3 codon lengths
        fresh.load_fs_profile(3, model.rfv(5), model.tfv(5))
    assert e.value.code == capi.EINVAL
    big_rfv = np.ones((367, 1101), np.float32)             # M = 1100 exceeds the single-warp kernels (32 lanes x 32 nodes)
    big_tfv = np.full((8, 1101), 0.1, np.float32)
    with pytest.raises(capi.BathGpuError) as e:
        fresh.load_fs_profile(3, big_rfv, big_tfv)
    assert e.value.code == capi.EINVAL and "1024" in str(e.value)
    fresh.load_fs_profile(5, model.rfv(5), model.tfv(5))
    env = capi.Context.make_windows([1], [300], nj=0.0)
    with pytest.raises(capi.BathGpuError) as e:            # trace buffer too small is an error, not a truncation
        fresh.fs_domains(env, max_steps=5)
    assert e.value.code == capi.EINVAL and "trace" in str(e.value)
    fresh.close()


@pytest.mark.parametrize("L", [3, 4, 5, 6, 7, 11, 12])
def test_shortest_windows(oracle, gpu_ctx, L):
    po = oracle
    from bath_b200 import capi
    model = po.Model(common.golden("2OG-FeII_Oxy_3.bhmm"))
    rng = np.random.default_rng(L)
    dsq = common.random_dna(rng, 64)
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    gpu_ctx.upload_block(dsq)
    starts = [1, 2, 3, 64 - L + 1]
    w = capi.Context.make_windows(starts, [L] * 4)
    sc, st = gpu_ctx.fs_fwd_windows(w)
    for t, s in enumerate(starts):
        ost, osc, _ = oracle_fwd(po, model, dsq, s, L)
        assert st[t] == ost
        if ost == 0:
            assert abs(sc[t] - osc) <= 1e-3, (L, s, sc[t], osc)
        else:
            assert np.isinf(sc[t]) or np.isnan(sc[t])


def test_all_degenerate_and_very_long_windows(oracle, gpu_ctx):
    po = oracle
    from bath_b200 import capi
    model = po.Model(common.golden("AMP_N.bhmm"))
    rng = np.random.default_rng(1)
    n = 30000
    dsq = common.random_dna(rng, n)
    dsq[1:601] = 15                                       # 600 N's
    dsq[5000:5003] = [5, 9, 11]                           # other degenerate codes (R, S, H)
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    gpu_ctx.upload_block(dsq)
    wins = [(1, 600), (300, 600), (4900, 300), (1, n), (7, n - 7)]       # all-N, half-N, mixed codes, the whole block
    w = capi.Context.make_windows([s for s, _ in wins], [l for _, l in wins])
    sc, st = gpu_ctx.fs_fwd_windows(w)
    for t, (s, L) in enumerate(wins):
        ost, osc, _ = oracle_fwd(po, model, dsq, s, L)
        assert st[t] == ost and abs(sc[t] - osc) <= 2e-3, (t, s, L, sc[t], osc)
    mocc, btot, etot, fsc, bsc, st = gpu_ctx.fs_bck_decode(w[:3], (0.5, 0.5), (0.9, 0.9, 0.9))
    assert np.all(st == 0) and np.all(np.abs(fsc - bsc) <= 1e-3)
    assert all(np.all(np.isfinite(m)) for m in mocc)


def test_tiny_and_degenerate_orfs(oracle, gpu_ctx):
    po = oracle
    lib = po.lib()
    from bath_b200 import capi
    from test_gpu_orf_filters import host_params
    model = po.Model(common.golden("AMP_N.bhmm"))
    rbv, rwv, twv, _, _ = model.om_tables()
    gpu_ctx.load_filter_profile(host_params(po, model, 16, 8), rbv, rwv, twv)
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    rng = np.random.default_rng(2)
    seqs = [np.array([7], np.uint8), np.array([26, 26, 26], np.uint8), np.full(40, 26, np.uint8),
            rng.integers(0, 20, 2).astype(np.uint8), rng.integers(0, 20, 2000).astype(np.uint8)]
    gpu_ctx.upload_orfs(np.concatenate(seqs))
    orfs = np.zeros(len(seqs), capi.orf_dtype)
    off = 0
    o = model.om.contents
    for t, s in enumerate(seqs):
        lib.bo_oprofile_ReconfigLength(model.om, len(s))
        orfs[t]["offset"], orfs[t]["L"], orfs[t]["tjb_b"], orfs[t]["xw_move"] = off, len(s), o.tjb_b, o.xw[1][0]
        off += len(s)
    msv, mst = gpu_ctx.msv_orfs(orfs)
    vit, vst, _ = gpu_ctx.vit_orfs(orfs)
    fwd, fst = gpu_ctx.fwd_orfs(orfs)
    for t, s in enumerate(seqs):
        d = np.concatenate([[255], s, [255]]).astype(np.uint8)
        lib.bo_oprofile_ReconfigLength(model.om, len(s))
        a, v, f = C.c_float(), C.c_float(), C.c_float()
        assert mst[t] == lib.bo_MSVFilter(po.u8ptr(d), len(s), model.om, C.byref(a)) and msv[t] == a.value
        assert vst[t] == lib.bo_ViterbiFilter(po.u8ptr(d), len(s), model.om, C.byref(v)) and vit[t] == v.value
        ost = lib.bo_ForwardParser(po.u8ptr(d), len(s), model.om, C.byref(f))
        assert fst[t] == ost and (ost != 0 or abs(fwd[t] - f.value) <= 1e-3)
    with pytest.raises(capi.BathGpuError):                # ORF outside the uploaded residues
        bad = orfs[:1].copy(); bad["offset"] = off
        gpu_ctx.msv_orfs(bad)


def test_slots_hold_two_targets(oracle, gpu_ctx):
    """the two target slots are independent: scores of a window depend only on the selected slot's block"""
    po = oracle
    from bath_b200 import capi
    model = po.Model(common.golden("AMP_N.bhmm"))
    rng = np.random.default_rng(3)
    a, b = common.random_dna(rng, 2000), common.random_dna(rng, 2000)
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    w = capi.Context.make_windows([1, 700], [600, 900])
    gpu_ctx.select_slot(0); gpu_ctx.upload_block(a)
    gpu_ctx.select_slot(1); gpu_ctx.upload_block(b)
    sb, _ = gpu_ctx.fs_fwd_windows(w)
    gpu_ctx.select_slot(0)
    sa, _ = gpu_ctx.fs_fwd_windows(w)
    for dsq, sc in ((a, sa), (b, sb)):
        for t in range(2):
            _, osc, _ = oracle_fwd(po, model, dsq, int(w["start"][t]), int(w["L"][t]))
            assert abs(sc[t] - osc) <= 1e-3
    assert not np.allclose(sa, sb)


def test_revcomp_slot_is_the_bottom_strand(oracle, gpu_ctx):
    """bathgpu_revcomp_slot: windows scored on the device-made reverse complement equal the oracle's on the host-made one
    (degenerate codes included: R<->Y, M<->K, ... as esl_sq_ReverseComplement maps them)"""
    po = oracle
    from bath_b200 import capi
    model = po.Model(common.golden("AMP_N.bhmm"))
    rng = np.random.default_rng(9)
    top = common.random_dna(rng, 5000)
    top[100:110] = [5, 6, 7, 8, 9, 10, 11, 12, 13, 14]
    bottom = po.revcomp_dsq(top)
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    gpu_ctx.select_slot(0); gpu_ctx.upload_block(top)
    gpu_ctx.revcomp_slot(0, 1)
    gpu_ctx.select_slot(1)
    w = capi.Context.make_windows([1, 1200, 4401, 4880], [900, 900, 600, 121])
    sc, st = gpu_ctx.fs_fwd_windows(w)
    gpu_ctx.select_slot(0)
    for t in range(len(w)):
        ost, osc, _ = oracle_fwd(po, model, bottom, int(w["start"][t]), int(w["L"][t]))
        assert st[t] == ost and abs(sc[t] - osc) <= 1e-3
    with pytest.raises(capi.BathGpuError):
        gpu_ctx.revcomp_slot(1, 1)
