"""scratch: Forward parser time against resident warps per SM (BATHGPU_FWD_WARPS) -- one process per setting (the cap is read once)"""
import os, subprocess, sys
code = r'''
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as po
from bath_b200 import capi
import common
ctx = capi.Context(0)
model = po.Model(common.golden("tRNA-synthetases.bhmm"), int(sys.argv[1]))
ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
rng = np.random.default_rng(1)
Lw, nwin = 1200, 16384
dsq = common.random_dna(rng, nwin * Lw)
ctx.upload_block(dsq)
w = capi.Context.make_windows(1 + np.arange(nwin) * Lw, np.full(nwin, Lw))
ctx.stage_windows(w)
for _ in range(2): ctx.fs_fwd_staged()
ms = min((ctx.fs_fwd_staged(), ctx.last_stage_timing()[0])[1] for _ in range(4))
import os
wps = int(os.environ.get("BATHGPU_FWD_WARPS", "0"))
cells = nwin * Lw * model.M
rows_per_warp = nwin * Lw / (148 * wps) if wps else 0
print(f"M={model.M} fwd={os.environ.get('BATHGPU_FWD','-')} warps/SM={wps:2d}: {ms:8.3f} ms {cells/ms/1e6:7.1f} GCUPS; cycles per row per warp = {ms*1e-3*1.91e9/rows_per_warp if wps else 0:7.1f}")
'''
for idx in (1,):
    for fwd in ("3", "4"):
        for wps in (4, 8, 12, 16):
            env = dict(os.environ, BATHGPU_FWD=fwd, BATHGPU_FWD_WARPS=str(wps))
            subprocess.run([sys.executable, "-c", code, str(idx)], env=env)
