import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, common, bench
from bath_b200 import capi, hostapi, synth
from oracle import pyoracle as po
models = [hostapi.QueryModel(common.golden("tRNA-synthetases.bhmm"), i) for i in range(3)]
rng = np.random.default_rng(11)
contigs, _ = synth.planted_contigs(rng, 6_000_000, [m.mat() for m in models], every=25_000, fs_rates=[m.fsprob for m in models], min_len=1_000_000, max_len=3_000_000)
ctxs = [capi.Context(0), capi.Context(0)]
be, keep = po.cpu_backend(16)
_, gt, gs, gh = bench.run_search(models, contigs, gpu_ctxs=ctxs)
_, ct, cs, ch = bench.run_search(models, contigs, backends=be)
for k in range(3):
    print(k, len(gh[k]), len(ch[k]), bench.compare_tables(gt[k], ct[k]))
    for a, b in zip(gt[k].splitlines(), ct[k].splitlines()):
        if a != b:
            print("G", a); print("C", b)
    for a, b in zip(gh[k], ch[k]):
        if abs(a["score"] - b["score"]) > 2e-3 or abs(a["bias"] - b["bias"]) > 2e-3:
            print("hit", a["name"], a["ali_from"], a["score"], b["score"], a["bias"], b["bias"], a["lnP"], b["lnP"])
