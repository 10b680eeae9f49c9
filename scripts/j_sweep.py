"""scratch: Forward / Forward+Backward GCUPS for every node count per lane the shipped models reach, with and without BATHGPU_J_BUMP"""
import os, subprocess, sys
code = r'''
import sys, numpy as np, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as po
from bath_b200 import capi
import common
ctx = capi.Context(0)
models = [("tRNA-proteins.bhmm", i) for i in range(12)] + [("2OG-FeII_Oxy_3.bhmm", 0), ("AMP_N.bhmm", 0), ("synthetic_M208.bhmm", 0)]
seen = set()
for hmm, idx in models:
    model = po.Model(common.golden(hmm), idx)
    if model.M in seen: continue
    seen.add(model.M)
    ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    rng = np.random.default_rng(1)
    Lw, nwin = 1200, 16384
    dsq = common.random_dna(rng, nwin * Lw)
    ctx.upload_block(dsq)
    w = capi.Context.make_windows(1 + np.arange(nwin) * Lw, np.full(nwin, Lw))
    ctx.stage_windows(w)
    for _ in range(2): ctx.fs_fwd_staged()
    ms = min((ctx.fs_fwd_staged(), ctx.last_stage_timing()[0])[1] for _ in range(3))
    nb = 8192
    ctx.fs_fwd_bck_xrows(w[:nb])
    tb = min((ctx.fs_fwd_bck_xrows(w[:nb]), ctx.last_stage_timing()[0])[1] for _ in range(2))
    print(f"bump={os.environ.get('BATHGPU_J_BUMP','0')} M={model.M:4d} J={(model.M+31)//32}: fwd {nwin*Lw*model.M/ms/1e6:7.1f}  fwd+bck {2*nb*Lw*model.M/tb/1e6:7.1f} GCUPS", flush=True)
'''
for bump in ("0", "1"):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BATHGPU_J_BUMP=bump))
