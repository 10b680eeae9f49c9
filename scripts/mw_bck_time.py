"""scratch: Forward (X rows) + Backward parser on long models, one-warp vs multi-warp Backward (BATHGPU_BCK_MW), one process per setting"""
import os, subprocess, sys
code = r'''
import sys, numpy as np, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as po
from bath_b200 import capi
import common
ctx = capi.Context(0)
for hmm, idx in [("MET-ct4.bhmm", 0), ("synthetic_M624.bhmm", 0), ("synthetic_M903.bhmm", 0)]:
    model = po.Model(common.golden(hmm), idx)
    ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    rng = np.random.default_rng(1)
    Lw, nwin = 1200, 4096
    dsq = common.random_dna(rng, nwin * Lw)
    ctx.upload_block(dsq)
    w = capi.Context.make_windows(1 + np.arange(nwin) * Lw, np.full(nwin, Lw))
    ctx.stage_windows(w)
    for _ in range(2): ctx.fs_fwd_staged()
    tf = min((ctx.fs_fwd_staged(), ctx.last_stage_timing()[0])[1] for _ in range(3))
    ctx.fs_fwd_bck_xrows(w)
    tb = min((ctx.fs_fwd_bck_xrows(w), ctx.last_stage_timing()[0])[1] for _ in range(3))
    c = nwin * Lw * model.M
    print(f"BCK_MW={os.environ.get('BATHGPU_BCK_MW','1')} M={model.M}: fwd {c/tf/1e6:6.0f}  fwd(xrows)+bck {2*c/tb/1e6:6.0f} GCUPS (both sweeps)  => bck alone ~{c/max(tb-tf,1e-9)/1e6:6.0f}", flush=True)
'''
for mw in ("0", "1", "2"):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BATHGPU_BCK_MW=mw))
