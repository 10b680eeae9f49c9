"""Forward parser on long models: multi-warp kernel (default) vs the one-warp kernel (BATHGPU_FWD_MW=0 in the environment)."""
import os, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common
from oracle import pyoracle as po          # reads the profile files; checks a sample of scores
from bath_b200 import capi
ctx = capi.Context(0)
rng = np.random.default_rng(42)
for hmm, idx in [m for m in [("synthetic_M624.bhmm", 0), ("synthetic_M903.bhmm", 0)] if os.environ.get("MW_ONLY", "") in m[0]]:
    model = po.Model(common.golden(hmm), idx)
    ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    for Lw in (1200,):
        nwin = 4096
        dsq = common.random_dna(rng, nwin * Lw)
        ctx.upload_block(dsq)
        w = capi.Context.make_windows(1 + np.arange(nwin) * Lw, np.full(nwin, Lw))
        ctx.stage_windows(w)
        for _ in range(2): ctx.fs_fwd_staged()
        t = min((ctx.fs_fwd_staged(), ctx.last_stage_timing()[0])[1] for _ in range(3))
        sc, st = ctx.fetch_scores(nwin)
        po.lib().bo_fs_oprofile_ReconfigLength(model.om_fs3, Lw // 3)
        ref, rst = po.batch_forward_parser(model, dsq, 1 + np.arange(64) * Lw, np.full(64, Lw), nthreads=16)
        print(f"MW={os.environ.get('BATHGPU_FWD_MW','1')} M={model.M} Lw={Lw} GCUPS={nwin * Lw * model.M / t / 1e6:.0f} max|gpu-oracle|={np.abs(sc[:64] - ref).max():.2e} status_ok={(st == 0).all()}", flush=True)
