"""scratch: why is the X-row stage slow for tRNA-synthetases[2] (M=247)?  kernel time on random windows vs windows over planted homologs"""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from bath_b200 import capi, hostapi, synth
HMM = 'tests/golden/tRNA-synthetases.bhmm'
ctx = capi.Context(0)
for idx in (1, 2):
    m = hostapi.QueryModel(HMM, idx)
    ctx.load_fs_profile(3, m.rfv(3), m.tfv(3))
    ctx.load_fs_profile(5, m.rfv(5), m.tfv(5))
    rng = np.random.default_rng(5)
    n = 20_000_000
    d, plants = synth.planted_genome(rng, n, m.mat(), every=20000, fs_rate=m.fsprob, revcomp_fraction=0.0)
    ctx.select_slot(0); ctx.upload_block(d)
    Lw = 1500
    hs = np.array([max(1, a - 200) for a, b, s in plants if a - 200 + Lw < n], np.int64)
    rs = hs + 8000                                        # iid stretches between the homologs
    for name, st in (("random", rs), ("homolog", hs)):
        for cnt in (150, len(st)):
            w = capi.Context.make_windows(st[:cnt], np.full(cnt, Lw))
            ctx.fs_fwd_bck_xrows(w)
            t0 = time.perf_counter(); ctx.fs_fwd_bck_xrows(w); wall = (time.perf_counter() - t0) * 1e3
            ms = ctx.last_stage_timing()[0]
            sc, stt = ctx.fs_fwd_windows(w); ms1 = ctx.last_stage_timing()[0]
            print(f"M={m.M} {name:8s} n={cnt:5d}: fwd+bck kernels {ms:8.3f} ms (wall {wall:8.2f}), fwd-only {ms1:8.3f} ms, max score {sc.max():.1f}, bad status {(stt != 0).sum()}", flush=True)
    # envelopes: the planted homologs themselves
    env = np.zeros(len(plants), capi.window_dtype)
    for z, (a, b, s) in enumerate(plants):
        L = b - a + 1
        env[z]["start"], env[z]["L"] = a, L
        pm = np.float32(2.0) / (np.float32(L // 3) + np.float32(2.0))
        env[z]["pmove"], env[z]["ploop"] = pm, np.float32(1.0) - pm
    for cnt in (110, len(env)):
        ctx.fs_domains(env[:cnt])
        t0 = time.perf_counter(); res, tr = ctx.fs_domains(env[:cnt]); wall = (time.perf_counter() - t0) * 1e3
        print(f"M={m.M} fs_domains n={cnt:5d}: kernels {ctx.last_stage_timing()[0]:8.3f} ms (wall {wall:8.2f}), bad status {(res['status'] != 0).sum()}", flush=True)
