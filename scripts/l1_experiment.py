"""scratch: does the Forward parser wait on table loads that miss L1?  Same kernel, same work, windows over a k-letter alphabet:
with 2 letters the codon rows touched are 4 + 8 + 16 of 336 (the whole table footprint 22 KB at M = 192), with 4 letters all of it."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import pyoracle as po
from bath_b200 import capi
import common
ctx = capi.Context(0)
for hmmfile, idx in [("AMP_N.bhmm", 0), ("tRNA-synthetases.bhmm", 1), ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0)]:
    model = po.Model(common.golden(hmmfile), idx)
    ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    for letters in (4, 3, 2, 1):
        rng = np.random.default_rng(1)
        Lw, nwin = 1200, 16384
        dsq = np.full(nwin * Lw + 2, 255, np.uint8)
        dsq[1:-1] = rng.integers(0, letters, nwin * Lw)
        ctx.upload_block(dsq)
        w = capi.Context.make_windows(1 + np.arange(nwin) * Lw, np.full(nwin, Lw))
        ctx.stage_windows(w)
        for _ in range(3): ctx.fs_fwd_staged()
        ts = []
        for _ in range(5):
            ctx.fs_fwd_staged(); ts.append(ctx.last_stage_timing()[0])
        ms = min(ts); cells = nwin * Lw * model.M
        rows = letters ** 2 + letters ** 3 + letters ** 4
        print(f"M={model.M} letters={letters} (table rows touched {rows:3d}, {rows * 4 * ((model.M + 31) // 32 * 32) / 1024:.0f} KB): {ms:.3f} ms  {cells/ms/1e6:.1f} GCUPS", flush=True)
