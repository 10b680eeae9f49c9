"""scratch: quick kernel timing on the GPU box (not the bench)"""
import sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import pyoracle as po
from bath_b200 import capi
sys.path.insert(0, 'tests')
import common
ctx = capi.Context(0)
print(ctx.device_info())
models = [("AMP_N.bhmm",0), ("tRNA-synthetases.bhmm",0), ("tRNA-synthetases.bhmm",2), ("PTHR37536.bhmm",0)]
if "--big" in sys.argv: models = [("MET-ct4.bhmm",0), ("synthetic_M624.bhmm",0), ("synthetic_M903.bhmm",0)]
for hmmfile, idx in models:
    model = po.Model(common.golden(hmmfile), idx)
    ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    rng = np.random.default_rng(1)
    for Lw in (1200,):
        nwin = 16384
        dsq = common.random_dna(rng, nwin * Lw)
        ctx.upload_block(dsq)
        w = capi.Context.make_windows(1 + np.arange(nwin) * Lw, np.full(nwin, Lw))
        ctx.stage_windows(w)
        for _ in range(3): ctx.fs_fwd_staged()
        ts = []
        for _ in range(5):
            ctx.fs_fwd_staged(); ts.append(ctx.last_stage_timing()[0])
        ms = min(ts); cells = nwin * Lw * model.M
        print(f"{hmmfile}[{idx}] M={model.M} Lw={Lw} nwin={nwin}: {ms:.3f} ms  {cells/ms/1e6:.1f} GCUPS  (all: {[round(t,3) for t in ts]})")
