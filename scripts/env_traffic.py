"""scratch: the per-envelope domain stage (bathgpu_fs_domains: 5-codon Forward, Backward + posterior decoding + null2, optimal accuracy,
traceback) over 2048 envelopes of 600 nt holding planted homologs, M = 192 -- run under ncu for the DRAM bytes per DP cell, or alone
for the stage time."""
import sys, time; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, common
from bath_b200 import capi
from oracle import pyoracle as po
ctx = capi.Context(0)
model = po.Model(common.golden("tRNA-synthetases.bhmm"), 1)
ctx.load_fs_profile(5, model.rfv(5), model.tfv(5))
nenv, L = 2048, 600
rng = np.random.default_rng(0)
dsq = common.random_dna(rng, nenv * L)
mat = common.hmm_mat(model)
for e in range(nenv):
    ins = common.sample_homolog(rng, mat, fs_rate=0.01, stop_rate=0.0)[:L - 20]
    dsq[1 + e * L + 10: 1 + e * L + 10 + len(ins)] = ins
ctx.upload_block(dsq)
env = capi.Context.make_windows(1 + np.arange(nenv) * L, np.full(nenv, L), nj=0.0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    t0 = time.perf_counter()
    res, tr = ctx.fs_domains(env, xfE5=(1.0, 0.0))
    dt = time.perf_counter() - t0
cells = nenv * L * model.M
print(f"{nenv} envelopes x {L} nt x M={model.M}: {cells/1e6:.1f} M cells, call {dt*1e3:.2f} ms, status ok {(res['status']==0).sum()}")
