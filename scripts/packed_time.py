"""scratch: upload and one-call Forward on a byte block against the same block host-packed"""
import sys, time; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, common
from bath_b200 import capi
from oracle import pyoracle as po
ctx = capi.Context(0)
model = po.Model(common.golden("tRNA-synthetases.bhmm"), 1)
ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
n = 100_000_000
rng = np.random.default_rng(0)
d = common.random_dna(rng, n)
dsq = capi.pinned_array(d.shape, np.uint8); dsq[:] = d
packed = capi.pinned_array(((n + 1) // 2,), np.uint8); capi.pack_dna4(dsq, out=packed)
nwin = n // 1200
wins = capi.pinned_array((nwin,), capi.window_dtype)
wins[:] = capi.Context.make_windows(1 + np.arange(nwin) * 1200, np.full(nwin, 1200))
sc = capi.pinned_array((nwin,), np.float32); st = capi.pinned_array((nwin,), np.int32)
def t(f, reps=8):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e3
print("upload bytes   %.3f ms" % t(lambda: ctx.upload_block(dsq)))
print("upload packed4 %.3f ms" % t(lambda: ctx.upload_block_packed4(packed, n)))
print("block bytes    %.3f ms" % t(lambda: ctx.fs_fwd_block_into(dsq, wins, (0.5, 0.5), sc, st)))
print("block packed4  %.3f ms" % t(lambda: ctx.fs_fwd_block_packed4_into(packed, n, wins, (0.5, 0.5), sc, st)))
print("block bytes    %.3f ms" % t(lambda: ctx.fs_fwd_block_into(dsq, wins, (0.5, 0.5), sc, st)))
print("block packed4  %.3f ms" % t(lambda: ctx.fs_fwd_block_packed4_into(packed, n, wins, (0.5, 0.5), sc, st)))
ctx.upload_block(dsq); ctx.stage_windows(wins)
print("staged kernel  %.3f ms" % t(lambda: ctx.fs_fwd_staged((0.5, 0.5))))
