"""The CPU arm's SIMD build of the 3-codon Forward parser (oracle/fwd3_avx2.c) against the scalar oracle it restates
(oracle/fs_fwdback.c = src/impl_sse/fwdback_fs.c:97-533): scores within 1e-4 nat (the products are fused in the SIMD build, the
scalar one rounds every operation), status codes equal -- random windows, frameshifted homologs that force rescaling, degenerate
nucleotides, ragged lengths, all shipped model sizes."""
import numpy as np
import pytest

import common
from test_gpu_fs_forward import make_block


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("2OG-FeII_Oxy_3.bhmm", 0), ("tRNA-synthetases.bhmm", 1),
                                           ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0), ("MET-ct4.bhmm", 0)])
def test_simd_forward_parser_matches_the_scalar_oracle(oracle, hmmfile, index):
    po = oracle
    if not po.simd_supported():
        pytest.skip("no AVX2 + FMA on this host")
    model = po.Model(common.golden(hmmfile), index)
    rng = np.random.default_rng(100 + index)
    dsq, wins = make_block(rng, model, n_random=12, n_homolog=20,
                           lengths=[60, 61, 62, 63, 300, 449, 600, 3 * max(model.max_length, 300) + 5])
    starts, lengths = [s for s, _ in wins], [l for _, l in wins]
    want, wst = po.batch_forward_parser(model, dsq, starts, lengths, 2)
    got, gst = po.batch_forward_parser(model, dsq, starts, lengths, 3, simd=True)
    assert np.array_equal(gst, wst)
    ok = wst == 0
    assert ok.sum() >= len(wins) - 2
    assert np.max(np.abs(got[ok] - want[ok])) <= 1e-4, np.max(np.abs(got[ok] - want[ok]))
    assert want[ok].max() > 20.0            # homologs are in there, and long ones rescale


def test_simd_forward_parser_short_and_degenerate_windows(oracle):
    po = oracle
    if not po.simd_supported():
        pytest.skip("no AVX2 + FMA on this host")
    model = po.Model(common.golden("AMP_N.bhmm"), 0)
    rng = np.random.default_rng(5)
    dsq = common.random_dna(rng, 4000, p_degenerate=0.05)
    lengths = list(range(3, 40)) + [100, 333]
    starts = [int(rng.integers(1, 4000 - l)) for l in lengths]
    want, wst = po.batch_forward_parser(model, dsq, starts, lengths, 1)
    got, gst = po.batch_forward_parser(model, dsq, starts, lengths, 2, simd=True)
    assert np.array_equal(gst, wst)
    ok = wst == 0
    assert np.allclose(got[ok], want[ok], atol=1e-4)
    assert np.array_equal(np.isinf(got[~ok]), np.isinf(want[~ok]))


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("2OG-FeII_Oxy_3.bhmm", 0), ("tRNA-synthetases.bhmm", 0), ("tRNA-synthetases.bhmm", 1),
                                           ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0), ("MET-ct4.bhmm", 0)])
def test_simd_msv_filter_is_the_scalar_one_bit_for_bit(oracle, hmmfile, index):
    """oracle/msv_avx2.c (the CPU arm's MSV + SSV screen) == bo_MSVFilter: status and score of random ORFs, ORFs with X residues,
    whole / partial / two-segment / diverged homologs (SSV answers, the J-state recursion, overflows), every ORF length's tjb cost"""
    import ctypes as C
    from test_gpu_orf_filters import make_orfs
    po, lib = oracle, oracle.lib()
    if not lib.bo_msv_simd_supported():
        pytest.skip("no AVX2 on this host")
    model = po.Model(common.golden(hmmfile), index)
    rng = np.random.default_rng(50 + index)
    seqs = make_orfs(rng, model, n_random=120, n_homolog=60)
    seqs += [rng.integers(0, 20, n).astype(np.uint8) for n in (1, 2, 3, 19, 20, 31, 32, 33, 1500)]
    lib.bo_msv_simd_create.restype = C.c_void_p
    lib.bo_msv_simd_create.argtypes = [C.c_void_p]
    lib.bo_msv_simd_destroy.argtypes = [C.c_void_p]
    lib.bo_MSVFilter_simd.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_int, C.c_void_p, C.POINTER(C.c_float)]
    im = lib.bo_msv_simd_create(model.om)
    assert im
    n_ssv_no = n_overflow = n_high = 0
    for s in seqs:
        L = len(s)
        d = np.concatenate([[255], s, [255]]).astype(np.uint8)
        lib.bo_oprofile_ReconfigLength(model.om, L)
        a, b = C.c_float(), C.c_float()
        st0 = lib.bo_MSVFilter(po.u8ptr(d), L, model.om, C.byref(a))
        st1 = lib.bo_MSVFilter_simd(im, po.u8ptr(d), L, model.om, C.byref(b))
        assert st0 == st1, (L, st0, st1)
        assert a.value == b.value or (np.isinf(a.value) and np.isinf(b.value)), (L, a.value, b.value)
        c = C.c_float()
        n_ssv_no += lib.bo_SSVFilter(po.u8ptr(d), L, model.om, C.byref(c)) != 0
        n_overflow += st0 == 16
        n_high += st0 == 0 and a.value > 5.0
    lib.bo_msv_simd_destroy(im)
    assert n_high + n_overflow >= 10 and n_ssv_no >= 5, (n_high, n_ssv_no, n_overflow)      # the SSV answer, the J-state recursion and overflows are exercised


def test_cpu_backend_with_simd_screen_writes_the_same_table(oracle):
    """the host pipeline over the CPU stage calls with the AVX2 MSV screen (bench.py's CPU arm) == with the scalar one"""
    from bath_b200 import hostapi, synth
    if not oracle.lib().bo_msv_simd_supported():
        pytest.skip("no AVX2 on this host")
    model = hostapi.QueryModel(common.golden("tRNA-synthetases.bhmm"), 1)
    rng = np.random.default_rng(3)
    dsq, _ = synth.planted_genome(rng, 600_000, model.mat(), every=20_000, fs_rate=model.fsprob)
    out = []
    for simd in (False, True):
        be, keep = oracle.cpu_backend(4, simd=simd)
        s = hostapi.Search(model, backend=be)
        s.queue_sequence("contig1", dsq)
        s.finish()
        out.append((s.tblout(), s.stats()["pos_past_msv"], s.stats()["n_hits_reported"]))
        s.close()
        del keep
    assert out[0] == out[1] and out[0][2] >= 20
