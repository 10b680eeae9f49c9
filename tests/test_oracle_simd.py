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
