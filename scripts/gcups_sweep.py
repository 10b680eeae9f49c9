"""Forward / Backward parser throughput over window length and model length (BASELINE.json config 5): device-resident inputs,
CUDA-event time of the kernels (bathgpu_last_stage_timing), iid-ACGT windows, 16 384 windows per point (fewer for the longest)."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common
from oracle import pyoracle as po          # only to read the profile files into the load calls' tables
from bath_b200 import capi

ctx = capi.Context(0)
models = [("AMP_N.bhmm", 0), ("tRNA-synthetases.bhmm", 0), ("tRNA-synthetases.bhmm", 1), ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0),
          ("MET-ct4.bhmm", 0), ("synthetic_M624.bhmm", 0), ("synthetic_M903.bhmm", 0)]
lengths = [300, 600, 1200, 2400, 4800]
rng = np.random.default_rng(42)
print("| M | Lw | windows | Forward GCUPS | Forward + Backward (X rows kept) GCUPS |")
print("|---|---|---|---|---|")
for hmm, idx in models:
    model = po.Model(common.golden(hmm), idx)
    ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    for Lw in lengths:
        nwin = 16384 if Lw * model.M < 1.5e6 else 4096
        dsq = common.random_dna(rng, nwin * Lw)
        ctx.upload_block(dsq)
        w = capi.Context.make_windows(1 + np.arange(nwin) * Lw, np.full(nwin, Lw))
        ctx.stage_windows(w)
        for _ in range(2): ctx.fs_fwd_staged()
        t = min((ctx.fs_fwd_staged(), ctx.last_stage_timing()[0])[1] for _ in range(3))
        cells = nwin * Lw * model.M
        nb = min(nwin, 8192 if Lw <= 1200 else 4096)      # several waves of windows per resident warp, and X rows that still fit in host memory
        ctx.fs_fwd_bck_xrows(w[:nb])
        tb = min((ctx.fs_fwd_bck_xrows(w[:nb]), ctx.last_stage_timing()[0])[1] for _ in range(2))
        print(f"| {model.M} | {Lw} | {nwin} | {cells / t / 1e6:.0f} | {2 * nb * Lw * model.M / tb / 1e6:.0f} |", flush=True)
