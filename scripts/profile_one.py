"""scratch: one config for ncu"""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as po
from bath_b200 import capi
import common
hmmfile, idx, Lw, nwin = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ctx = capi.Context(0)
model = po.Model(common.golden(hmmfile), idx)
ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
rng = np.random.default_rng(1)
dsq = common.random_dna(rng, nwin * Lw)
ctx.upload_block(dsq)
w = capi.Context.make_windows(1 + np.arange(nwin) * Lw, np.full(nwin, Lw))
ctx.stage_windows(w)
for _ in range(4): ctx.fs_fwd_staged()
print(ctx.last_stage_timing())
