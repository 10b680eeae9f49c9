"""scratch: Forward parser generations at 12 and 16 nodes per lane (one process per setting: the choice is read once)"""
import os, subprocess, sys
code = r'''
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as po
from bath_b200 import capi
import common, os
ctx = capi.Context(0)
for hmm, idx in [("synthetic_M377.bhmm", 0), ("MET-ct4.bhmm", 0), ("MET-ct4.bhmm", 1)]:
    model = po.Model(common.golden(hmm), idx)
    ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    rng = np.random.default_rng(1)
    Lw, nwin = 1200, 16384
    dsq = common.random_dna(rng, nwin * Lw)
    ctx.upload_block(dsq)
    w = capi.Context.make_windows(1 + np.arange(nwin) * Lw, np.full(nwin, Lw))
    ctx.stage_windows(w)
    for _ in range(2): ctx.fs_fwd_staged()
    ms = min((ctx.fs_fwd_staged(), ctx.last_stage_timing()[0])[1] for _ in range(4))
    print(f"FWD={os.environ.get('BATHGPU_FWD','auto')} MW={os.environ.get('BATHGPU_FWD_MW','1')} M={model.M}: {ms:8.3f} ms {nwin*Lw*model.M/ms/1e6:7.1f} GCUPS", flush=True)
'''
for fwd, mw in (("3", "1"), ("4", "1"), ("3", "2")):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BATHGPU_FWD=fwd, BATHGPU_FWD_MW=mw))
