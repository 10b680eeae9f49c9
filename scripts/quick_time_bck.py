"""scratch: Forward (X rows kept) + Backward parser timing on the GPU box (not the bench)"""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as po
from bath_b200 import capi
import common
ctx = capi.Context(0)
for hmmfile, idx in [("AMP_N.bhmm", 0), ("tRNA-synthetases.bhmm", 1), ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0)]:
    model = po.Model(common.golden(hmmfile), idx)
    ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    rng = np.random.default_rng(1)
    Lw, nwin = 1200, 8192
    dsq = common.random_dna(rng, nwin * Lw)
    ctx.upload_block(dsq)
    w = capi.Context.make_windows(1 + np.arange(nwin) * Lw, np.full(nwin, Lw))
    ctx.fs_fwd_bck_xrows(w)
    ts = []
    for _ in range(3):
        ctx.fs_fwd_bck_xrows(w); ts.append(ctx.last_stage_timing()[0])
    ms = min(ts)
    print(f"{hmmfile}[{idx}] M={model.M}: fwd+bck {ms:.3f} ms  {2 * nwin * Lw * model.M / ms / 1e6:.1f} GCUPS (both sweeps)")
